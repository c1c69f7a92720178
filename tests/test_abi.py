"""The C-ABI library loads, exports every symbol include/shn.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import golden_io
import hnsw_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "shn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(shn_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(pkg):
    lib = pkg.shn.lib()
    syms = declared_symbols()
    assert "shn_search" in syms and "shn_index_load" in syms
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/shn.h but not exported by libshn_b200.so"


def test_header_is_plain_c():
    src = '#include "shn.h"\nint main(void) { return (int)sizeof(shn_stats) == 0; }\n'
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                   input=src.encode(), check=True)


def test_library_is_sm100a_sass(pkg):
    out = subprocess.run(["cuobjdump", "-lelf", pkg.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_stats_struct_layout(pkg):
    assert C.sizeof(pkg.shn.Stats) == 10 * 8 + 3 * 8 + 4 * 8


def test_argument_errors_need_no_gpu(pkg):
    lib = pkg.shn.lib()
    assert lib.shn_set_option(None, b"warps_per_sm", 4) == -1
    assert b"null" in lib.shn_last_error()
    assert lib.shn_index_size(None) == 0
    lib.shn_index_free(None)


def test_no_cpu_fallback(pkg):
    """Without a usable sm_100 device every compute entry point returns SHN_ERR_CUDA; nothing runs on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    case = golden_io.load_case("l2_d8_n40_m32")
    with pytest.raises(pkg.ShnError) as e:
        pkg.Index.from_dumps(case["dumps"], case["dim"], case["m"])
    assert e.value.code == -3


def test_malformed_dump_is_rejected(pkg):
    case = golden_io.load_case("l2_d8_n40_m32")
    with pytest.raises(pkg.ShnError) as e:  # wrong dim: records do not tile the file
        pkg.repartition_dumps(case["dumps"], case["dim"] + 4, case["m"], 1)
    assert e.value.code == -2
    with pytest.raises(pkg.ShnError):
        pkg.repartition_dumps([case["dumps"][0][:100]], case["dim"], case["m"], 1)


@pytest.mark.parametrize("name,parts", [("l2_d32_n2000_m16", 3), ("l2_d96_n1500_m16_2mn", 1), ("l2_d20_n300_m4_3mn", 2),
                                        ("ip_d40_n1500_m8", 1)])
def test_dump_writer_round_trip(pkg, name, parts):
    """Host code only: parse reference dumps, re-emit them for another memory-node count, and let the oracle search
    the rewritten dumps — same ids, same distance bits, same counters as on the reference's own dumps."""
    case = golden_io.load_case(name)
    out = pkg.repartition_dumps(case["dumps"], case["dim"], case["m"], parts)
    assert len(out) == parts
    for d in out:
        assert int(np.frombuffer(d[:8], np.uint64)[0]) == d.size  # free_ptr == file size (memory_node.hh:187-195)
    a = hnsw_oracle.Index(case["dumps"], case["dim"], case["m"])
    b = hnsw_oracle.Index([d.tobytes() for d in out], case["dim"], case["m"])
    if parts == len(case["dumps"]) == 1:
        # same size, same graph (the reference leaves stale pointers in unused list slots, so not the same bytes)
        assert out[0].size == len(case["dumps"][0])
        ea, eb = a.export(), b.export()
        for key in ea:
            assert (ea[key] == eb[key]).all(), key
    (k, ef) = next(iter(case["runs"]))
    ra = a.knn(case["queries"], k, ef, ip=case["ip"], counters=True)
    rb = b.knn(case["queries"], k, ef, ip=case["ip"], counters=True)
    assert (ra[0] == rb[0]).all() and (ra[1].view(np.uint32) == rb[1].view(np.uint32)).all()
    assert (ra[3]["distcomps"] == rb[3]["distcomps"]).all()


@pytest.mark.parametrize("name", ["l2_d32_n2000_m16", "ip_d40_n1500_m8", "l2_d20_n300_m4_3mn"])
def test_level_recipe_matches_the_reference_build(pkg, name):
    """Host code only: the levels the GPU builder will draw equal the levels in a dump the reference itself built with
    the same seed and m (fixtures: seed 1234, one thread, four coroutines — the first inserts race for the empty index
    and are written at level 0, hnsw.hh:56-85, so they are left out of the comparison)."""
    case = golden_io.load_case(name)
    ex = hnsw_oracle.Index(case["dumps"], case["dim"], case["m"]).export()
    ref = ex["level"][np.argsort(ex["uid"])]
    got = pkg.draw_levels(case["n"], case["m"], seed=1234)
    assert got[0] == 0 and len(got) == case["n"]
    assert (got[4:] == ref[4:]).all()
    assert ref.max() >= 1


def test_level_recipe_is_capped_and_seeded(pkg):
    a = pkg.draw_levels(100000, 2, seed=5)
    assert (pkg.draw_levels(100000, 2, seed=5) == a).all() and not (pkg.draw_levels(100000, 2, seed=6) == a).all()
    top = np.maximum.accumulate(a)
    assert (np.diff(top) <= 1).all()  # never more than one above the current top (hnsw.hh:106)


def test_compact_visited_key_is_a_bijection():
    """search.cuh visited_compact stores 15 bits of mix24(id) in bucket (mix24(id) >> 15): exact only if mix24 is a bijection of
    the 24-bit integers.  The constants are read from the CUDA source and the whole domain is checked."""
    import re
    import numpy as np
    src = open(os.path.join(ROOT, "dm-hnsw-reference_b200", "csrc", "search.cuh")).read()
    body = src[src.index("uint32_t mix24(uint32_t h)"):]
    body = body[:body.index("return h;")]
    steps = re.findall(r"h = \(h \* (0x[0-9A-Fa-f]+)u\) & 0xFFFFFFu; h \^= h >> (\d+);", body)
    assert len(steps) == 3, body
    h = np.arange(1 << 24, dtype=np.uint64)
    for mul, shift in steps:
        assert int(mul, 16) % 2 == 1 and 0 < int(shift) < 24
        h = (h * np.uint64(int(mul, 16))) & np.uint64(0xFFFFFF)
        h ^= h >> np.uint64(int(shift))
    assert len(np.unique(h)) == 1 << 24
    # and the buckets are evenly used: 512 buckets, 32768 ids each
    assert (np.bincount((h >> np.uint64(15)).astype(np.int64), minlength=512) == 32768).all()
