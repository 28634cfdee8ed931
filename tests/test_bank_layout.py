"""Shared-memory bank arithmetic of the wide-block kernels, checked on the host (no GPU).

The kernels rely on three layouts being conflict-free (csrc/k1_wide.cuh produce(), csrc/k1_large.cuh plane_skew /
nf_row_share); ncu showed 31 % of k1_wide's shared-memory wavefronts to be two-way conflicts before they were introduced.
A 32-bit access is served per warp, a 64-bit access per 16 lanes; a wavefront touches each of the 32 banks once.
"""
import itertools

HW_PITCH = 273      # complex slots per consume column (k1_large.cuh)
PITCH_P = 280       # complex slots per produce column (k1_wide.cuh K1W_PITCH_P)


def wavefronts_64(word_addrs):
    """64-bit accesses of 16 lanes (word address of the low half): number of wavefronts = max lanes per bank pair"""
    per = {}
    for a in word_addrs:
        assert a % 2 == 0
        per.setdefault((a // 2) % 16, set()).add(a)
    return max(len(v) for v in per.values())


def wavefronts_32(word_addrs):
    per = {}
    for a in word_addrs:
        per.setdefault(a % 32, set()).add(a)
    return max(len(v) for v in per.values())


def produce_lane(lane, warp):
    hlp = (lane & 7) | ((lane >> 4) << 3)
    fp = 2 * warp + ((lane >> 3) & 1)
    return hlp, fp


def test_produce_lane_assignment_covers_both_columns():
    for warp in range(8):
        got = sorted(produce_lane(lane, warp) for lane in range(32))
        assert got == sorted(itertools.product(range(16), (2 * warp, 2 * warp + 1)))


def test_swizzled_tile_accesses_are_conflict_free_per_half_warp():
    """element (row hlp + 16 j, column fp) of a 128B-swizzled [256][16] float2 tile: one wavefront per 16 lanes"""
    for warp in range(8):
        for j in range(16):
            for half in range(2):
                addrs = []
                for lane in range(16 * half, 16 * half + 16):
                    hlp, fp = produce_lane(lane, warp)
                    byte = hlp * 128 + (((fp >> 1) ^ (hlp & 7)) << 4) + ((fp & 1) << 3) + j * 2048
                    addrs.append(byte // 4)
                assert wavefronts_64(addrs) == 1
    # the plain assignment (one column per half-warp) is the two-way conflict ncu reported
    addrs = [(hl * 128 + (((0 >> 1) ^ (hl & 7)) << 4)) // 4 for hl in range(16)]
    assert wavefronts_64(addrs) == 2


def test_produce_transposes_are_conflict_free_with_pitch_8_mod_16():
    assert PITCH_P % 16 == 8 and PITCH_P >= 16 * 17
    for warp in range(8):
        for k in range(16):
            for half in range(2):
                st, ld = [], []
                for lane in range(16 * half, 16 * half + 16):
                    hlp, fp = produce_lane(lane, warp)
                    st.append(2 * (fp * PITCH_P + k * 17 + hlp))   # col[k * 17 + hl] = v
                    ld.append(2 * (fp * PITCH_P + hlp * 17 + k))   # v = col[hl * 17 + k]
                assert wavefronts_64(st) == 1 and wavefronts_64(ld) == 1


def plane_skew(row_word0, j):
    return (2 * (j >> 1) + 16 * (j & 1) - row_word0) & 31


def wide_plane_of(j):
    col = (j >> 1) * (2 * PITCH_P) + (j & 1) * HW_PITCH
    return 2 * col + plane_skew(2 * col, j)


def mid8k_plane_of(r):
    w0 = r * 2 * HW_PITCH
    return w0 + plane_skew(w0, r >> 1)


def test_planes_stay_inside_their_own_column():
    for j in range(16):
        col = 2 * ((j >> 1) * (2 * PITCH_P) + (j & 1) * HW_PITCH)
        assert col <= wide_plane_of(j) and wide_plane_of(j) + 512 <= col + 2 * HW_PITCH
    for r in range(32):
        assert r * 2 * HW_PITCH <= mid8k_plane_of(r) and mid8k_plane_of(r) + 256 <= (r + 1) * 2 * HW_PITCH


def test_plane_stores_take_one_wavefront_per_warp():
    """prow[k2s] = psd: lanes (hl, row f) of a warp, rows 2w and 2w + 1, sixteen consecutive positions each"""
    for w in range(8):
        for base in (0, 128, 240):
            addrs = [wide_plane_of(2 * w + (lane >> 4)) + base + (lane & 15) for lane in range(32)]
            assert wavefronts_32(addrs) == 1
    for h in range(2):
        for wg in range(8):
            addrs = [mid8k_plane_of(2 * (2 * wg + (lane >> 4)) + h) + 64 + (lane & 15) for lane in range(32)]
            assert wavefronts_32(addrs) == 1


def test_window_share_loads_take_one_wavefront_per_warp():
    """sixteen rows of one window per half-warp from a common start; the upper half rotated when the starts have equal parity"""
    for p_lo, p_hi in itertools.product(range(0, 40, 3), range(20, 70, 5)):
        rot = 1 if (p_hi - p_lo) % 2 == 0 else 0
        for step in range(1, 24):
            addrs = [wide_plane_of(lane & 15) + (p_lo + step if lane < 16 else p_hi + rot + step) for lane in range(32)]
            assert wavefronts_32(addrs) == 1
            for h in range(2):
                addrs = [mid8k_plane_of(2 * (lane & 15) + h) + (p_lo + step if lane < 16 else p_hi + rot + step) for lane in range(32)]
                assert wavefronts_32(addrs) == 1
