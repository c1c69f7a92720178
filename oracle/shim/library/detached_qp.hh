// oracle/shim: see library/connection_manager.hh.  TEST INFRASTRUCTURE ONLY.
// Surface used by src/shared_context.hh:17-21 (ctor with delegated CQs, connect, .qp).
#pragma once
#include <library/connection_manager.hh>

class DetachedQP {
public:
  DetachedQP(Context& context, ibv_cq* send_cq, ibv_cq* recv_cq)
      : qp(std::make_unique<QueuePair>(&context, send_cq, recv_cq)) {}
  void connect(Context&, u16, QP&) const {}
  QP qp;
};
