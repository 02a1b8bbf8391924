"""Index algebra of the tensor-core kernels (csrc/pinnjet_tc.cuh: TcGeo, TcThread, tc_load_owner, tc_store_rows,
tc_reduce_points; csrc/pinnjet_k1tc3.cuh / pinnjet_k2tc2.cuh), restated in numpy.

A tile is 128 GEMM rows r = CP*p + c (point p, channel c, channels padded to CP in {2, 4, 8}) x 64 hidden units.  The kernels
move such a block between three layouts: the TMEM row layout (lane = row, one tcgen05.ld.32x32b.x16 per warp = 32 rows x 16
units), the K-major SWIZZLE_128B shared-memory images the MMAs read, and the OWNER layout of the epilogues (a thread holds
all channels of one point x UG adjacent units) -- which is also the layout of the z-jet records K1-TC leaves for K2-TC.
The formulas below are the ones in the kernels; the tests pin their invariants."""
import numpy as np
import pytest

ROWS, H, NCW = 128, 64, 16
STAGE_STRIDE = 20


def geo(C):
    CP = 2 if C <= 2 else (4 if C <= 4 else 8)
    PW = 32 // CP
    NUG = 32 // PW
    return dict(C=C, CP=CP, TP=ROWS // CP, PW=PW, NUG=NUG, UG=16 // NUG, REC=C * (16 // NUG))


def thread(C, tid):
    g = geo(C)
    warp, lane = tid >> 5, tid & 31
    q, j = warp & 3, (warp >> 2) & 3
    pt, ug = lane // g["NUG"], lane % g["NUG"]
    p = q * g["PW"] + pt
    return dict(g, warp=warp, lane=lane, q=q, j=j, pt=pt, ug=ug, p=p, ubase=16 * j + g["UG"] * ug, R0=g["CP"] * p)


def sw128_off(row, chunk16):
    return (row >> 3) * 1024 + (row & 7) * 128 + ((chunk16 ^ (row & 7)) << 4)


def img_off(t, c):   # TcThread::img_off
    row_off = (t["R0"] >> 3) * 1024 + (t["R0"] & 7) * 128
    r7 = (t["R0"] + c) & 7
    return row_off + c * 128 + ((((t["ubase"] >> 3) ^ r7) << 4) + (t["ubase"] & 7) * 2)


@pytest.mark.parametrize("C", [1, 2, 3, 4, 5, 6, 7])
def test_owner_layout_is_a_partition_of_the_tile(C):
    g = geo(C)
    assert g["UG"] * 2 == g["PW"] and g["TP"] * g["CP"] == ROWS
    seen = np.zeros((g["TP"], H), dtype=int)
    for tid in range(NCW * 32):
        t = thread(C, tid)
        seen[t["p"], t["ubase"]:t["ubase"] + t["UG"]] += 1
    assert np.all(seen == 1)                       # every (point, unit) has exactly one owner thread


@pytest.mark.parametrize("C", [2, 4, 5])
def test_tmem_block_to_owner_layout_through_the_private_staging_block(C):
    """tc_load_owner: warp (q, j) reads rows 32q.. x units 16j.. (lane = row), stores its 16 values at stage[lane][0:16],
    and reads back rows CP*pt + c, columns UG*ug ..: exactly the (point, channel, unit) values it owns."""
    g = geo(C)
    acc = np.arange(ROWS * H, dtype=np.float64).reshape(ROWS, H)      # accumulator [row][unit]
    for warp in range(NCW):
        q, j = warp & 3, (warp >> 2) & 3
        stage = np.full((32, STAGE_STRIDE), np.nan)
        for lane in range(32):
            stage[lane, :16] = acc[32 * q + lane, 16 * j:16 * j + 16]
        for lane in range(32):
            t = thread(C, warp * 32 + lane)
            for c in range(C):
                got = stage[g["CP"] * t["pt"] + c, g["UG"] * t["ug"]:g["UG"] * t["ug"] + g["UG"]]
                want = acc[t["R0"] + c, t["ubase"]:t["ubase"] + g["UG"]]
                np.testing.assert_array_equal(got, want)
    # conflict-free 16-byte accesses: the 8 lanes of a quarter-warp hit 8 distinct 16-byte bank groups
    for lane0 in range(0, 32, 8):
        assert len({((lane0 + k) * STAGE_STRIDE * 4 // 16) % 8 for k in range(8)}) == 8


@pytest.mark.parametrize("C", [2, 3, 4, 5, 7])
def test_image_rows_written_by_owner_threads_tile_the_swizzled_image(C):
    g = geo(C)
    owner = {}
    for tid in range(NCW * 32):
        t = thread(C, tid)
        for c in range(C):
            off = img_off(t, c)
            row, u = t["R0"] + c, t["ubase"]
            assert off == sw128_off(row, (u * 2) >> 4) + ((u * 2) & 15)          # tc_store_rows == the MMA's K-major layout
            for b in range(off, off + 2 * g["UG"]):
                assert b not in owner
                owner[b] = tid
    assert len(owner) == (ROWS // g["CP"]) * C * H * 2                             # padded channel rows stay untouched (zero)


@pytest.mark.parametrize("C", [2, 4, 5])
def test_reduce_scatter_over_the_point_lanes(C):
    """tc_reduce_points: UG values summed over the PW point lanes with UG shuffles; lane pt ends with value pt >> 1."""
    g = geo(C)
    PW, NUG, UG = g["PW"], g["NUG"], g["UG"]
    rng = np.random.default_rng(C)
    v = rng.normal(size=(32, UG))
    lanes = np.arange(32)
    pt = lanes // NUG
    w, cnt, bit = v.copy(), UG, PW // 2
    while cnt > 1:
        up = (pt & bit) != 0
        new = w.copy()
        for i in range(cnt // 2):
            keep = np.where(up, w[:, i + cnt // 2], w[:, i])
            send = np.where(up, w[:, i], w[:, i + cnt // 2])
            new[:, i] = keep + send[lanes ^ (bit * NUG)]
        w, cnt, bit = new, cnt // 2, bit // 2
    res = w[:, 0] + w[lanes ^ NUG, 0]
    for lane in lanes:
        ug = lane % NUG
        want = sum(v[p * NUG + ug, pt[lane] >> 1] for p in range(PW))
        assert abs(res[lane] - want) < 1e-12
    # the adders (even pt) of a warp write distinct units: no atomics needed inside a warp
    units = [(lane % NUG) * UG + (pt[lane] >> 1) for lane in lanes if pt[lane] % 2 == 0]
    assert len(set(units)) == len(units) == 16


@pytest.mark.parametrize("C", [2, 4, 5])
def test_record_blocks_are_indexed_by_thread(C):
    g = geo(C)
    assert g["REC"] * 4 % 8 == 0 and (g["UG"] < 4 or g["REC"] * 4 % 16 == 0)       # float2 / float4 accesses stay aligned
    assert NCW * 32 * g["REC"] == g["TP"] * C * H                                 # one block = every (point, channel, unit) once
    assert (NCW * 32 * g["REC"] * 4) % 16 == 0                                     # bulk-TMA size
