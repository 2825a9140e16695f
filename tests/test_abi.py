"""CPU tier: the C-ABI library loads and exports exactly what include/rfnet_ops.h declares (no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rfnet_ops.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = re.findall(r"\b(rfnet_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S)
    return {name: [a.strip() for a in args.split(",") if a.strip() and a.strip() != "void"] for name, args in decls}


def test_header_declares_every_reference_launcher():
    names = set(declared_functions())
    # one entry point per launcher of the reference + workspace queries + host variants
    for required in ["rfnet_nn_distance", "rfnet_nn_distance_grad", "rfnet_approxmatch", "rfnet_matchcost", "rfnet_matchcostgrad",
                     "rfnet_farthestpointsampling", "rfnet_gatherpoint", "rfnet_scatteraddpoint", "rfnet_query_ball_point", "rfnet_group_point",
                     "rfnet_group_point_grad", "rfnet_three_nn", "rfnet_three_interpolate", "rfnet_three_interpolate_grad",
                     "rfnet_nn_distance_host", "rfnet_emd_host"]:
        assert required in names, required


def test_library_exports_every_declared_symbol():
    from rfnet_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from rfnet_b200 import build
        build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    decl = declared_functions()
    assert len(decl) >= 24
    for name in decl:
        assert hasattr(lib, name), "librfnet_ops.so does not export " + name
    exported = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (rfnet_[a-z0-9_]+)", exported))
    assert exported == set(decl), "header and library disagree: %s" % (exported ^ set(decl))


def test_ctypes_signatures_match_header_arity():
    from rfnet_b200 import _lib
    decl = declared_functions()
    assert set(_lib.SIGNATURES) == set(decl)
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        assert len(argtypes) == len(decl[name]), "%s: ctypes has %d args, header %d" % (name, len(argtypes), len(decl[name]))
        for ct, text in zip(argtypes, decl[name]):
            is_ptr = "*" in text or "rfnet_stream_t" in text
            assert (ct is ctypes.c_void_p) == is_ptr, "%s: %s" % (name, text)
            if "size_t" in text and not is_ptr:
                assert ct is ctypes.c_size_t


def test_library_has_sm100a_code_and_no_other_arch():
    from rfnet_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_version_and_error_string_without_gpu():
    from rfnet_b200 import _lib
    lib = _lib.load()
    assert lib.rfnet_version() >= 100
    assert b"invalid argument" in lib.rfnet_error_string(1)
    # workspace queries are pure host arithmetic
    # packed keys (8 B per point, rounded to 16 B) + prepared candidate rows (16 B per point, clouds padded to whole groups of 16)
    # + 16 B of origin / norm per possible work-item range (256 candidates at least)
    def expect(b, n, m):
        pad = lambda c: (c + 15) // 16 * 16
        return (8 * b * (n + m) + 15) // 16 * 16 + 16 * b * (pad(n) + pad(m)) + 16 * b * ((n + 255) // 256 + (m + 255) // 256)
    for shape in ((32, 2048, 16384), (1, 10, 17), (3, 1000, 777)):
        assert lib.rfnet_nn_distance_workspace_bytes(*shape) == expect(*shape), shape
    assert lib.rfnet_approxmatch_workspace_bytes(2, 100, 100) > 0
    assert lib.rfnet_approxmatch_workspace_bytes(0, 100, 100) == 0


def test_nn_distance_launch_plans_cover_the_clouds_and_fit_the_workspace():
    """Host logic of the Chamfer search (no GPU): for a sweep of shapes the plan's work items tile the candidate range exactly
    once, chunks are whole groups of the kernel that scans them, an item of the filtered search never exceeds what one
    preparation CTA covers (4096 candidates), and the number of item ranges stays within the per-range records the workspace
    reserves (one per 256 candidates)."""
    import ctypes
    from rfnet_b200 import _lib
    lib = _lib.load()
    out = (ctypes.c_int * 10)()
    shapes = [(b, n, m) for b in (1, 3, 4, 32, 64, 257) for n in (1, 5, 300, 1024, 2048, 2049, 5000, 16384, 40000) for m in (7, 1000, 2048, 4100, 16384, 65536)]
    seen_filter = seen_direct = 0
    for b, n, m in shapes:
        for flags in (0, 2):
            assert lib.rfnet_nn_distance_plan(b, n, m, flags, out) == 0
            direct, q = out[0], out[1]
            assert q in (2, 4, 8)
            assert direct or flags == 0
            seen_direct += direct
            seen_filter += 1 - direct
            for d, (nq, nc) in enumerate(((n, m), (m, n))):
                chunk, cps, nsplit, items = out[2 + 4 * d], out[3 + 4 * d], out[4 + 4 * d], out[5 + 4 * d]
                assert 0 < chunk <= 1024 and chunk % (8 if direct else 16) == 0
                nchunks = -(-nc // chunk)
                assert 1 <= cps <= nchunks and nsplit == -(-nchunks // cps), (b, n, m, flags, d)
                assert items == b * -(-nq // (128 * q)) * nsplit
                if not direct:
                    assert cps * chunk <= 4096, (b, n, m, d)
                    assert nsplit <= -(-nc // 256), (b, n, m, d)           # per-range records in the workspace
    assert seen_filter > 50 and seen_direct > 50
