"""Index algebra of the tensor-core kernels (pinnjet_k1tc2.cuh, pinnjet_k2tc.cuh), restated in numpy.

The kernels move a [rows x units] block between three layouts: the TMEM row layout (one thread per GEMM row), the K-major
SWIZZLE_128B shared-memory images the MMAs read, and the "owner" layout of the epilogue (a thread holds all channels of two
adjacent points x UG adjacent units).  The formulas below are the ones in the kernels; the test pins their invariants:
staging round trip, exclusive ownership of image chunks per warp, and that the hoisted address form equals sw128_off."""
import numpy as np
import pytest

IMG = 256 * 128   # bytes of one split image of the forward kernel (256 rows x 64 bf16)


def sw128_off(row, chunk16):
    return (row >> 3) * 1024 + (row & 7) * 128 + ((chunk16 ^ (row & 7)) << 4)


def warp_geometry(C, warp, lane):
    PW = 32 // C
    NPP = PW // 2
    NUG = 32 // NPP
    UG = 32 // NUG
    hf, g, q = warp >> 3, (warp >> 2) & 1, warp & 3
    rowbase = g * 128 + q * 32
    ppidx, ug = lane // NUG, lane % NUG
    return dict(UG=UG, NUG=NUG, hf=hf, rowbase=rowbase, ppidx=ppidx, ug=ug, R0=rowbase + 2 * C * ppidx,
                ubase=hf * 32 + ug * UG)


@pytest.mark.parametrize("C", [2, 4])
def test_forward_epilogue_staging_and_image_addresses(C):
    D = np.arange(256 * 64, dtype=np.float64).reshape(256, 64)        # TMEM accumulators: D[row][unit]
    smem = np.full(3 * IMG // 4, -1.0)                                # images as 4-byte slots
    owner_of_chunk = {}
    for warp in range(16):
        for lane in range(32):
            geo = warp_geometry(C, warp, lane)
            myrow, hf = geo["rowbase"] + lane, geo["hf"]
            stage_wr = (myrow >> 3) * 1024 + (myrow & 7) * 128
            stage_c = (hf * 4) ^ (myrow & 7)
            for s in range(8):                                        # TMEM row -> staging chunks (kernel form)
                off = stage_wr + (s >> 2) * IMG + ((stage_c ^ (s & 3)) << 4)
                assert off == (s >> 2) * IMG + sw128_off(myrow, hf * 4 + (s & 3))
                assert owner_of_chunk.setdefault(off, warp) == warp   # no two warps share a staging chunk
                assert np.all(smem[off // 4: off // 4 + 4] == -1)
                smem[off // 4: off // 4 + 4] = D[myrow, hf * 32 + 4 * s: hf * 32 + 4 * s + 4]
    covered = np.zeros((256, 64), dtype=int)
    for warp in range(16):
        for lane in range(32):
            geo = warp_geometry(C, warp, lane)
            UG, hf, ug, R0 = geo["UG"], geo["hf"], geo["ug"], geo["R0"]
            own_row = (R0 >> 3) * 1024 + (R0 & 7) * 128
            awr_c = ((hf * 4 + ((ug >> 1) if UG == 4 else ug)) ^ (R0 & 7)) << 4
            awr_b = (ug & 1) * 8 if UG == 4 else 0
            for pp in range(2):
                for c in range(C):
                    j = C * pp + c
                    r = geo["rowbase"] + C * (2 * geo["ppidx"] + pp) + c
                    assert r == R0 + j
                    got = []
                    for s4 in range(UG // 4):                         # staging -> owner layout (kernel form)
                        s = ug * (UG // 4) + s4
                        sc = ((hf * 4 + (s & 3)) ^ (R0 & 7)) << 4
                        off = own_row + (s >> 2) * IMG + j * 128 + (sc ^ (j << 4))
                        assert owner_of_chunk[off] == warp            # only chunks parked by the own warp are read
                        got += list(smem[off // 4: off // 4 + 4])
                    assert np.array_equal(got, D[r, geo["ubase"]: geo["ubase"] + UG])
                    dst = own_row + j * 128 + ((awr_c ^ (j << 4)) + awr_b)   # A-row store of the next layer
                    ref = sw128_off(r, geo["ubase"] >> 3) + (geo["ubase"] & 7) * 2
                    assert dst == ref
                    covered[r, geo["ubase"]: geo["ubase"] + UG] += 1
    assert np.all(covered == 1)                                       # every (row, unit) is written exactly once


def test_m64_accumulator_lane_layout_and_mn_major_reading():
    """Facts measured by experiments/tcgen05_probe/probe_wgrad.cu: an M = 64 accumulator keeps row m in TMEM lane
    (m % 16) + 32 * (m / 16); a K-major SWIZZLE_128B image read MN-major advances 2048 B per K = 16 instruction."""
    lanes = [(m % 16) + 32 * (m // 16) for m in range(64)]
    assert len(set(lanes)) == 64 and max(lanes) == 111
    for q in range(4):                                                # warp q reads rows 16q .. 16q+15 in its first 16 lanes
        assert lanes[16 * q: 16 * q + 16] == list(range(32 * q, 32 * q + 16))
    # element (row r, unit u) of an image: the MN-major reader takes K = r, MN = u; 16 K-steps = 16 rows = 2 groups of 8
    for r in range(0, 128, 16):
        assert sw128_off(r, 0) == (r // 16) * 2048
