"""CPU checks of the packed-byte arithmetic k_fast_cells relies on (jetson_slam_b200/csrc/jsfe_kernels.cuh), restated in numpy
uint32 arithmetic: the kernel itself only runs on the GPU (tests/test_gpu_parity.py), but its tricks are pure integer identities
and can be proved exhaustively here.
  * phase A: on 6-bit pixels q = p >> 2, "p - v > t" implies "q_p - q_v >= m" with m = (t-3)//4 + 1 (0 for t < 3), and the byte lane
    q_p - q_v + (128 - m) never leaves [0, 255] (no carry or borrow between the four pixels of a word);
  * phase B: the per-byte compare msb_gt, the complement-by-multiply used for the dark polarity, the merge of the 16 ring flags
    into one index and the permuted LUT bit order built by jsfe_create."""
import numpy as np

U = np.uint32
M32 = 0xFFFFFFFF


def msb_gt(a, b):
    """per-byte MSB = (a > b), as in the kernel (unsigned bytes)"""
    a, b = a.astype(np.uint64), b.astype(np.uint64)
    nb7 = (~b) & 0x7F7F7F7F
    t = ((a & 0x7F7F7F7F) + nb7) & M32
    return (((a & ~b) | (~(a ^ b) & t)) & M32).astype(U)


def bitsel(a, b, mask):
    return ((a & ~U(mask)) | (b & U(mask))).astype(U)


def test_phase_a_filter_is_conservative_and_lane_safe():
    p, v = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    dq = (p >> 2) - (v >> 2)
    for t in range(256):
        m = (t - 3) // 4 + 1 if t >= 3 else 0
        assert 0 <= m <= 64
        bright, dark = (p - v) > t, (v - p) > t
        assert (dq[bright] >= m).all() and ((-dq)[dark] >= m).all(), t
        lane_b, lane_d = dq + 128 - m, -dq + 128 - m
        assert lane_b.min() >= 0 and lane_b.max() <= 255 and lane_d.min() >= 0 and lane_d.max() <= 255
        assert ((lane_b >= 128) == (dq >= m)).all()      # the byte MSB is the flag


def test_packed_lanes_do_not_interact():
    rng = np.random.default_rng(0)
    for t in (0, 7, 20, 60, 255):
        m = (t - 3) // 4 + 1 if t >= 3 else 0
        C4 = U((128 - m) * 0x01010101)
        P = rng.integers(0, 1 << 32, size=20000, dtype=np.uint64).astype(U)
        V = rng.integers(0, 1 << 32, size=20000, dtype=np.uint64).astype(U)
        QP, QV = (P >> U(2)) & U(0x3F3F3F3F), (V >> U(2)) & U(0x3F3F3F3F)
        fb = (QP + (C4 - QV)).astype(U)
        fd = ((C4 + QV) - QP).astype(U)
        for k in range(4):
            qp, qv = ((QP >> U(8 * k)) & U(0xFF)).astype(int), ((QV >> U(8 * k)) & U(0xFF)).astype(int)
            assert ((((fb >> U(8 * k + 7)) & U(1)) == 1) == (qp - qv >= m)).all()
            assert ((((fd >> U(8 * k + 7)) & U(1)) == 1) == (qv - qp >= m)).all()


def test_msb_gt_exhaustive_per_byte():
    a, b = np.meshgrid(np.arange(256, dtype=np.uint64), np.arange(256, dtype=np.uint64), indexing="ij")
    # put the pair in byte 1 with noisy neighbours in bytes 0 and 2
    wa = ((a << 8) | 0x00FF00A5).astype(U).ravel()
    wb = ((b << 8) | 0x00C3005A).astype(U).ravel()
    r = msb_gt(wa, wb)
    assert ((((r >> U(15)) & U(1)) == 1) == (a > b).ravel()).all()


def _index_of_mask(m):
    """jsfe_create's permutation (jsfe.cu): ring k = byte k%4 of word k/4 -> index bit 4*(k%4) + 3 - k/4"""
    idx = 0
    for k in range(16):
        if m & (1 << k):
            idx |= 1 << (4 * (k & 3) + 3 - (k >> 2))
    return idx


def test_flag_merge_gives_the_permuted_lut_index():
    rng = np.random.default_rng(1)
    n = 5000
    ring = rng.integers(0, 256, size=(n, 16), dtype=np.uint64)       # ring point k of candidate i
    v = rng.integers(0, 256, size=n)
    for t in (5, 20, 40):
        for dark in (False, True):
            R = [(ring[:, 4 * j] | (ring[:, 4 * j + 1] << 8) | (ring[:, 4 * j + 2] << 16) | (ring[:, 4 * j + 3] << 24)).astype(U) for j in range(4)]
            sgn, off = (U(M32), U(M32)) if dark else (U(1), U(0))
            X = [(r.astype(np.uint64) * int(sgn) + int(off)).astype(np.uint64) & M32 for r in R]     # complement by multiply-add
            X = [x.astype(U) for x in X]
            if dark:
                assert all((x == ~r).all() for x, r in zip(X, R))
            hi, lo = np.minimum(v + t, 255), np.maximum(v - t, 0)
            H4 = ((255 - lo if dark else hi).astype(np.uint64) * 0x01010101).astype(U)
            f = [msb_gt(x, H4) for x in X]
            W = bitsel(f[1] >> U(1), f[0], 0x80808080)
            W = bitsel(f[2] >> U(2), W, 0xC0C0C0C0)
            W = bitsel(f[3] >> U(3), W, 0xE0E0E0E0)
            u = bitsel(W >> U(8), W >> U(4), 0x0F0F0F0F)
            idx = (u & U(0xFF)) | ((u >> U(8)) & U(0xFF00))
            # naive: the ring mask of this polarity, then the host's permutation
            rp = ring.astype(int)
            flags = (rp < (v - t)[:, None]) if dark else (rp > (v + t)[:, None])
            want = np.array([_index_of_mask(int(sum(1 << k for k in range(16) if row[k]))) for row in flags], dtype=U)
            assert np.array_equal(idx, want), (t, dark)


def test_permutation_is_a_bijection():
    seen = {_index_of_mask(1 << k) for k in range(16)}
    assert seen == {1 << k for k in range(16)}


def test_positives_written_over_consumed_entries_never_clobber_an_unread_one():
    """Phase B of k_fast_cells keeps no separate positives array: warp w evaluates entries [256 it + 32 w, +32) in pass `it` and writes
    its q-th hit over the (q & 31)-th entry of its (q >> 5)-th pass (bright entries grow from the front of the list, dark ones from
    the back).  Model of that index arithmetic: every overwritten slot belongs to the writing warp and has been read, and phase C's
    walk (q-th positive of warp w at entry 256 (q >> 5) + 32 w + (q & 31)) finds exactly the warp's hits in order."""
    rng = np.random.default_rng(5)
    for trial in range(400):
        cap = int(rng.integers(1, 3000))
        ntot = int(rng.integers(0, cap + 1))
        nb = int(rng.integers(0, ntot + 1))
        density = rng.choice([0.0, 0.1, 0.4, 0.9, 1.0])
        hit = rng.random(ntot) < density

        def slot_of(j):
            return j if j < nb else cap - 1 - (j - nb)

        lst = [None] * cap
        for j in range(ntot):
            assert lst[slot_of(j)] is None          # bright and dark entries do not overlap while nb + nd <= cap
            lst[slot_of(j)] = ("entry", j)
        read = set()
        nit = (ntot + 255) >> 8
        wpos = [0] * 8
        hits_of = [[] for _ in range(8)]
        # warps of a block run concurrently: any interleaving is legal as long as a warp-pass reads all its entries before it writes
        order = [(it, w) for it in range(nit) for w in range(8)]
        if trial % 2:                               # arbitrary interleaving of the warps; only a warp's own passes stay in order
            nxt = [0] * 8
            order = []
            while len(order) < 8 * nit:
                w = int(rng.choice([x for x in range(8) if nxt[x] < nit]))
                order.append((nxt[w], w))
                nxt[w] += 1
        for it, w in order:
            mine = [256 * it + 32 * w + lane for lane in range(32)]
            codes = {}
            for i in mine:
                if i < ntot:
                    kind, j = lst[slot_of(i)]
                    assert kind == "entry" and j == i, "an entry was clobbered before it was read"
                    read.add(i)
                    codes[i] = j
            for i in mine:
                if i < ntot and hit[i]:
                    q = wpos[w]
                    wpos[w] += 1
                    jd = ((q >> 5) << 8) + 32 * w + (q & 31)
                    assert jd in read and (jd >> 5) & 7 == w and (jd >> 8) <= it
                    lst[slot_of(jd)] = ("pos", codes[i])
                    hits_of[w].append(i)
        for w in range(8):
            got = []
            for q in range(wpos[w]):
                kind, j = lst[slot_of(((q >> 5) << 8) + 32 * w + (q & 31))]
                assert kind == "pos"
                got.append(j)
            assert got == hits_of[w]
        assert sum(wpos) == int(hit.sum())


def test_emission_slots_stay_inside_the_list_and_never_overlap_when_it_fits():
    """Emission of k_fast_cells: every warp reserves (bright, dark) slot counts with ONE packed atomic add (bright | dark << 16),
    threads take consecutive slots in lane order, bright entries grow from slot 0, dark ones from slot cap - 1 downwards, and a
    thread whose entries would cross the end of the array writes nothing (the counters still record them, so the block knows the
    list overflowed and falls back).  Model: all stores are inside [0, cap); if bright + dark <= cap no slot is written twice and
    exactly bright + dark slots are written."""
    rng = np.random.default_rng(11)
    for trial in range(300):
        cap = int(rng.integers(1, 4000))
        n_chunks = int(rng.integers(1, 12))              # emission rounds of the block (8 warps each)
        density = rng.choice([0.02, 0.2, 0.6, 1.0])
        packed = 0                                        # s_ncand
        written = []
        for _ in range(n_chunks):
            for w in rng.permutation(8):                  # the warps' atomics land in any order
                cb = rng.binomial(64, density * rng.random(), size=32)     # per-lane counts (<= 64 pixels per thread and polarity)
                cd = rng.binomial(64, density * rng.random(), size=32)
                inc = np.cumsum(cb) | (np.cumsum(cd) << 16)                # packed inclusive scan (no carry: totals < 65536)
                tot = int(inc[31])
                base = packed
                packed += tot
                for lane in range(32):
                    ob = (base & 0xFFFF) + (int(inc[lane]) & 0xFFFF) - int(cb[lane])
                    od = (base >> 16) + (int(inc[lane]) >> 16) - int(cd[lane])
                    if ob + cb[lane] <= cap:
                        written += list(range(ob, ob + int(cb[lane])))
                    if od + cd[lane] <= cap:
                        written += [cap - 1 - k for k in range(od, od + int(cd[lane]))]
            if (packed & 0xFFFF) > 60000 or (packed >> 16) > 60000:
                break                                      # the kernel's counters are 16 bits: a block has at most 9216 x 2 positions
        nb, nd = packed & 0xFFFF, packed >> 16
        assert all(0 <= s < cap for s in written)
        if nb + nd <= cap:
            assert len(written) == nb + nd and len(set(written)) == len(written)
