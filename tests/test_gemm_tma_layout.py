"""CPU replay of the shared-memory layouts of csrc/gemm_tma.cu: where the TMA unit puts each operand element (128-byte
swizzle for k-contiguous tiles, 64-byte swizzle for 8-column blocks of m/n-contiguous tiles), which element every MMA
lane reads (the rho / nu permutations), and where every accumulator lands in C.  Checks (a) the product is exact for all
four storage forms and every tile configuration the launcher uses, (b) each half warp's sixteen 64-bit fragment reads hit
sixteen distinct 8-byte slots of a 128-byte bank window (no bank conflict)."""
import numpy as np
import pytest

BK = 16


def rho(g):
    return 2 * (g & 3) + (g >> 2)


def nu(g):
    return (g & 1) + 4 * ((g >> 1) & 1) + 2 * (g >> 2)


def place_kcontig(tile):
    """tile [R, 16] -> shared memory (doubles): row pitch 128 B, 16-byte chunk index ^= row & 7 (CU_TENSOR_MAP_SWIZZLE_128B)."""
    s = np.zeros(tile.shape[0] * 16)
    for r in range(tile.shape[0]):
        for k in range(16):
            s[r * 16 + (((k >> 1) ^ (r & 7)) << 1) + (k & 1)] = tile[r, k]
    return s


def place_mncontig(tile):
    """tile [16 k, C] -> one {8 columns, 16 k} box per 8-column block, 64-byte rows, chunk ^= (k >> 1) & 3 (SWIZZLE_64B)."""
    C = tile.shape[1]
    s = np.zeros(16 * C)
    for j in range(C // 8):
        for k in range(16):
            for c in range(8):
                s[j * 128 + k * 8 + (((c >> 1) ^ ((k >> 1) & 3)) << 1) + (c & 1)] = tile[k, 8 * j + c]
    return s


def replay(BM, BN, WGN, TA, TB, rng):
    WM, WN = BM // 2, BN // WGN
    MT, NT = WM // 8, WN // 8
    A, B = rng.standard_normal((BM, BK)), rng.standard_normal((BK, BN))
    sA = place_mncontig(A.T.copy()) if TA else place_kcontig(A)
    sB = place_kcontig(B.T.copy()) if TB else place_mncontig(B)
    C = np.zeros((BM, BN))
    conflicts = 0
    for warp in range(2 * WGN):
        wr, wc = warp // WGN, warp % WGN          # a warp's 8-row / 8-column blocks are interleaved over the CTA tile
        acc = np.zeros((32, MT, NT, 2))
        for kk in range(0, 16, 4):
            af, bf = np.zeros((32, MT)), np.zeros((32, NT))
            aaddr, baddr = np.zeros((32, MT), int), np.zeros((32, NT), int)
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                r, v = rho(g), nu(g)
                y2, z2 = 2 * ((t >> 1) ^ r), 2 * ((v >> 1) ^ (t >> 1))
                a_base = wr * 128 + t * 8 + (v & 1) if TA else (8 * wr + r) * 16 + (t & 1)
                b_base = (8 * wc + r) * 16 + (t & 1) if TB else wc * 128 + t * 8 + (v & 1)
                ak = kk * 8 + (z2 ^ (kk & 4)) if TA else (kk ^ y2)
                bk = (kk ^ y2) if TB else kk * 8 + (z2 ^ (kk & 4))
                for i in range(MT):
                    aaddr[lane, i] = a_base + i * 256 + ak
                    af[lane, i] = sA[aaddr[lane, i]]
                for j in range(NT):
                    baddr[lane, j] = b_base + j * WGN * 128 + bk
                    bf[lane, j] = sB[baddr[lane, j]]
            for arr in (aaddr, baddr):
                for col in range(arr.shape[1]):
                    for h in (0, 16):
                        conflicts += 16 - len(set(arr[h:h + 16, col] % 16))
            for i in range(MT):            # mma.m8n8k4: lane (g, t) supplies A[g][t], B[t][g]; holds D[g][2t], D[g][2t+1]
                for j in range(NT):
                    Af, Bf = np.zeros((8, 4)), np.zeros((4, 8))
                    for lane in range(32):
                        Af[lane >> 2, lane & 3] = af[lane, i]
                        Bf[lane & 3, lane >> 2] = bf[lane, j]
                    D = Af @ Bf
                    for lane in range(32):
                        acc[lane, i, j, 0] += D[lane >> 2, 2 * (lane & 3)]
                        acc[lane, i, j, 1] += D[lane >> 2, 2 * (lane & 3) + 1]
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            rg = nu(g) if TA else rho(g)
            c0l = rho(2 * t) if TB else nu(2 * t)
            c1l = rho(2 * t + 1) if TB else nu(2 * t + 1)
            for i in range(MT):
                for j in range(NT):
                    C[8 * (2 * i + wr) + rg, 8 * (j * WGN + wc) + c0l] = acc[lane, i, j, 0]
                    C[8 * (2 * i + wr) + rg, 8 * (j * WGN + wc) + c1l] = acc[lane, i, j, 1]
    return np.abs(C - A @ B).max(), conflicts


@pytest.mark.parametrize('cfg', [(80, 64, 2), (80, 80, 2), (80, 64, 4), (128, 128, 4), (128, 64, 4)])
@pytest.mark.parametrize('TA', [False, True])
@pytest.mark.parametrize('TB', [False, True])
def test_tma_tile_layout_is_exact_and_conflict_free(cfg, TA, TB):
    err, conflicts = replay(*cfg, TA, TB, np.random.default_rng(7))
    assert err < 1e-13
    assert conflicts == 0


def test_adjacent_accumulator_columns_for_vector_stores():
    # B stored [K, N]: the two accumulator columns of a lane are adjacent and even-aligned -> one 16-byte store
    for t in range(4):
        assert nu(2 * t + 1) == nu(2 * t) + 1 and nu(2 * t) % 2 == 0
    assert sorted(rho(g) for g in range(8)) == list(range(8)) and sorted(nu(g) for g in range(8)) == list(range(8))
