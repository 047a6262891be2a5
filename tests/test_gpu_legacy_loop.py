"""GPU: the rest of the reference's inner-loop boundary exported by libmp3gpu.so with the reference's own signatures
(include/mp3gpu_legacy.h; loop-pvt.h:46-117, loop.c:51-53): inner_loop, bin_search_StepSize, calc_runlen, count1_bitcount,
subdivide, bigv_tab_select, new_choose_table, bigv_bitcount.  Each is called side by side with the UNMODIFIED reference's
function (oracle/_ref/libref.so) on the same arguments — including caller-supplied gr_info fields that do not belong to the
spectrum, since the functions must use what they are given — and every gr_info field, return value and quantised value
must be equal."""
import ctypes as C

import numpy as np
import pytest

import oracle
import ref_harness

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_harness.have_ref(), reason="oracle/_ref/libref.so not built (reference sources absent)")]

FIELDS = ["big_values", "count1", "count1table_select", "region0_count", "region1_count", "table_select0", "table_select1",
          "table_select2", "address1", "address2", "address3"]


def libs(pkg):
    G = ref_harness.GrInfo
    vp, gp = C.c_void_p, C.POINTER(G)
    out = []
    for lib in (pkg.load_library(), C.CDLL(ref_harness.LIBREF)):
        lib.inner_loop.argtypes = [vp, vp, C.c_int, gp, C.c_int, C.c_int]
        lib.inner_loop.restype = C.c_int
        lib.bin_search_StepSize.argtypes = [C.c_int, C.c_double, vp, vp, gp]
        lib.bin_search_StepSize.restype = C.c_int
        lib.calc_runlen.argtypes = [vp, gp]
        lib.count1_bitcount.argtypes = [vp, gp]
        lib.count1_bitcount.restype = C.c_int
        lib.subdivide.argtypes = [gp]
        lib.bigv_tab_select.argtypes = [vp, gp]
        lib.new_choose_table.argtypes = [vp, C.c_uint, C.c_uint]
        lib.new_choose_table.restype = C.c_int
        lib.bigv_bitcount.argtypes = [vp, gp]
        lib.bigv_bitcount.restype = C.c_int
        out.append(lib)
    return out


def row(g):
    r = ref_harness.gr_to_row(g)
    return [r[ref_harness.GR_FIELDS.index(k)] for k in FIELDS] + [g.quantizerStepSize]


def fresh(bt, **kw):
    g = ref_harness.GrInfo()
    g.window_switching_flag = 1 if bt else 0
    g.block_type = bt
    for k, v in kw.items():
        setattr(g, k, v)
    return g


def spectra(golden, n, seed):
    g = golden["cfg3_48k_stereo_320"]
    ixs = np.ascontiguousarray(g["ix"][:, :, :2]).reshape(-1, 576).astype(np.int32)
    rng = np.random.default_rng(seed)
    sel = rng.integers(0, len(ixs), n)
    out = ixs[sel].copy()
    out[0] = 0                                   # all zero
    out[1] = 0; out[1, 575] = 1                  # a single one at the very end
    out[2] = 1                                   # everything count1
    out[3, :20] = 40                             # ESC values
    return out, rng


def test_count_pipeline_functions(pkg, golden):
    mine, ref = libs(pkg)
    ixs, rng = spectra(golden, 60, 11)
    for t, ix in enumerate(ixs):
        for bt in (0, 1, 2, 3):
            a, b = fresh(bt), fresh(bt)
            ix = np.ascontiguousarray(ix)
            for lib, g in ((mine, a), (ref, b)):
                lib.calc_runlen(ix.ctypes.data, C.byref(g))
            assert row(a) == row(b), ("calc_runlen", t, bt, row(a), row(b))
            ra, rb = mine.count1_bitcount(ix.ctypes.data, C.byref(a)), ref.count1_bitcount(ix.ctypes.data, C.byref(b))
            assert ra == rb and row(a) == row(b), ("count1_bitcount", t, bt, ra, rb)
            mine.subdivide(C.byref(a)); ref.subdivide(C.byref(b))
            assert row(a) == row(b), ("subdivide", t, bt, row(a), row(b))
            mine.bigv_tab_select(ix.ctypes.data, C.byref(a)); ref.bigv_tab_select(ix.ctypes.data, C.byref(b))
            assert row(a) == row(b), ("bigv_tab_select", t, bt, row(a), row(b))
            ra, rb = mine.bigv_bitcount(ix.ctypes.data, C.byref(a)), ref.bigv_bitcount(ix.ctypes.data, C.byref(b))
            assert ra == rb, ("bigv_bitcount", t, bt, ra, rb)


def test_functions_use_the_callers_fields(pkg, golden):
    """subdivide for every big_values; count1 / table selection / bit count with regions and tables that are NOT the
    spectrum's own"""
    mine, ref = libs(pkg)
    for bv in range(0, 289):
        for bt in (0, 1, 2):
            a, b = fresh(bt, big_values=bv, address1=4, address2=8, address3=12), fresh(bt, big_values=bv, address1=4, address2=8, address3=12)
            mine.subdivide(C.byref(a)); ref.subdivide(C.byref(b))
            assert row(a) == row(b), (bv, bt, row(a), row(b))
    ixs, rng = spectra(golden, 40, 5)
    for t, ix in enumerate(ixs):
        ix = np.ascontiguousarray(ix)
        small = np.ascontiguousarray(np.minimum(ix, 1))
        bv, c1 = int(rng.integers(0, 200)), int(rng.integers(0, 40))
        a, b = fresh(0, big_values=bv, count1=c1), fresh(0, big_values=bv, count1=c1)
        assert mine.count1_bitcount(small.ctypes.data, C.byref(a)) == ref.count1_bitcount(small.ctypes.data, C.byref(b)) and row(a) == row(b)
        a1 = 2 * int(rng.integers(0, 100)); a2 = a1 + 2 * int(rng.integers(0, 100)); bvv = int(rng.integers(a2 // 2, 289))
        a, b = fresh(0, big_values=bvv, address1=a1, address2=a2, address3=2 * bvv), fresh(0, big_values=bvv, address1=a1, address2=a2, address3=2 * bvv)
        mine.bigv_tab_select(ix.ctypes.data, C.byref(a)); ref.bigv_tab_select(ix.ctypes.data, C.byref(b))
        assert row(a) == row(b), ("tab_select", t, a1, a2, bvv, row(a), row(b))
        # tables that cover the values but are not the cheapest: the bit count must follow them
        mx = int(ix.max())
        for tabs in ((15, 24, 31), (13, 15, 16)) if mx <= 15 else ((23, 31, 30),):
            for g in (a, b):
                g.table_select[0], g.table_select[1], g.table_select[2] = tabs
            assert mine.bigv_bitcount(ix.ctypes.data, C.byref(a)) == ref.bigv_bitcount(ix.ctypes.data, C.byref(b)), (t, tabs)
        lo = 2 * int(rng.integers(0, 200)); hi = lo + 2 * int(rng.integers(1, 80))
        hi = min(hi, 576)
        assert mine.new_choose_table(ix.ctypes.data, lo, hi) == ref.new_choose_table(ix.ctypes.data, lo, hi), (t, lo, hi)


def test_inner_loop_and_bin_search(pkg, golden):
    mine, ref = libs(pkg)
    g = golden["cfg1_44k_stereo_128"]
    o = oracle.encode_stream(g["pcm"], 44100, 128)
    rng = np.random.default_rng(17)
    for t in range(30):
        f = int(rng.integers(2, len(o)))
        xr = np.ascontiguousarray(o["xr"][f]).copy() * float(rng.choice([1.0, 20.0, 300.0]))       # [2][2][576]
        gr, ch, bt = int(rng.integers(0, 2)), int(rng.integers(0, 2)), int(rng.choice([0, 0, 1, 2, 3]))
        max_bits = int(rng.integers(200, 3000))
        q0 = int(rng.integers(-80, 10))
        a, b = fresh(bt, quantizerStepSize=float(q0)), fresh(bt, quantizerStepSize=float(q0))
        ia, ib = np.zeros((2, 2, 576), np.int32), np.zeros((2, 2, 576), np.int32)
        ra = mine.inner_loop(xr.ctypes.data, ia.ctypes.data, max_bits, C.byref(a), gr, ch)
        rb = ref.inner_loop(xr.ctypes.data, ib.ctypes.data, max_bits, C.byref(b), gr, ch)
        assert ra == rb and row(a) == row(b) and np.array_equal(ia, ib), ("inner_loop", t, ra, rb, row(a), row(b))
        a, b = fresh(bt), fresh(bt)
        xa = np.ascontiguousarray(np.abs(xr[gr, ch]))
        ja, jb = np.zeros(576, np.int32), np.zeros(576, np.int32)
        ra = mine.bin_search_StepSize(max_bits, float(q0), ja.ctypes.data, xa.ctypes.data, C.byref(a))
        rb = ref.bin_search_StepSize(max_bits, float(q0), jb.ctypes.data, xa.ctypes.data, C.byref(b))
        assert ra == rb and row(a) == row(b) and np.array_equal(ja, jb), ("bin_search", t, ra, rb, row(a), row(b))
