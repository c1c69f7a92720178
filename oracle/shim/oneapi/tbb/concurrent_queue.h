// oracle/shim: oneTBB's concurrent_queue is included by rdma-library/library/types.hh:4
// but never instantiated (the alias there points at moodycamel).  Empty on purpose.
#pragma once
