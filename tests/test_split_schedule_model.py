"""Index logic of the split K2 schedule (digiham_b200/csrc/demod.cu: demod_search_kernel -> demod_volume_kernel ->
demod_slice_kernel) against the one-kernel walk (demod_kernel), on abstract values: every symbol is identified by the
absolute position of its window, every slicing decision by the set of symbols in the 100-entry volume ring
(reference src/gfsk_demodulator/gfsk_demodulator.cpp:18-122).  The model restates what the kernels index, not their
arithmetic; it pins the invariants the host-side sizing relies on:
  * blocks per call <= (carry_cap + n) / (100 sps - 1) + 1                     (split_blocks)
  * window starts grow with the symbol number b * 100 + j, also across blocks; a warp's 32 windows span
    <= 32 sps + 2 samples and end inside the visible stream                    (vol_warp_floats, RowView::stage)
  * symbol j of block b is output b * 100 + j - j_done_in, every output is written exactly once
  * exactly one slice group per channel and call hands the volume ring to the next call
The GPU tests (tests/test_demod_split_gpu.py) compare the real kernels with the compiled reference."""
import random

B = 100   # VARIANCE_SYMBOLS == VOLUME_RB_SIZE (include/gfsk_demodulator.hpp:5-6)


def processable(T, P, vo, sps):
    if T - P < sps + 2:
        return 0
    room = T - P - vo - sps - 2
    return 1 + (min(B - 1, room // sps) if room >= sps else 0)


def search(abs_start, vo):
    """stand-in for the variance search: a deterministic nudge in {-1, 0, +1} per block position"""
    return (hash((abs_start, vo, 12345)) % 3) - 1


class OneKernel:
    def __init__(self, sps):
        self.sps, self.vol_prev, self.vo, self.j_done, self.tail, self.base = sps, [("z", i) for i in range(B)], 0, 0, [], 0

    def call(self, chunk):
        sps = self.sps
        stream = self.tail + chunk
        T, vo, j_done, P, out = len(stream), self.vo, self.j_done, 0, []
        pv = list(self.vol_prev)
        m = processable(T, P, vo, sps)
        while m > j_done:
            vol = [None] * B
            for j in range(m):
                vol[j] = ("v", self.base + P + j * sps + (vo if j else 0))
            full = m == B
            if full:
                vo_next = search(self.base + P, vo)
                P_next = P + B * sps + vo
                m_next = processable(T, P_next, vo_next, sps)
            for j in range(j_done, m):
                out.append((vol[j], frozenset(vol[:j + 1] + pv[j + 1:])))
            if not full:
                j_done = m
                break
            P, vo, j_done, m, pv = P_next, vo_next, 0, m_next, vol
        self.vol_prev, self.vo, self.j_done = pv, vo, j_done
        self.tail = stream[P:]
        self.base += P
        return out


class Split:
    def __init__(self, sps):
        self.sps, self.tail, self.base = sps, [], 0
        self.state = dict(vol_prev=[("z", i) for i in range(B)], vo=0, j_done=0)

    def call(self, chunk):
        sps = self.sps
        stream = self.tail + chunk
        T = len(stream)
        st_in, st_out = self.state, dict(vol_prev=None, vo=None, j_done=None)
        # demod_search_kernel
        vo, j_done_in = st_in["vo"], st_in["j_done"]
        P, nfull, m, rec = 0, 0, processable(T, 0, st_in["vo"], sps), [(0, st_in["vo"])]
        while m == B:
            vo_next = search(self.base + P, vo)
            P_next = P + B * sps + vo
            if T - P_next < B * sps + 1:          # not pre-staged: the next block cannot be complete
                assert processable(T, P_next, vo_next, sps) < B
            nfull, P, vo, m = nfull + 1, P_next, vo_next, processable(T, P_next, vo_next, sps)
            rec.append((P, vo))
        j_cur = j_done_in if nfull == 0 else 0
        st_out["vo"], st_out["j_done"] = vo, (m if m > j_cur else j_cur)
        emitted = (m - j_done_in if m > j_done_in else 0) if nfull == 0 else (B - j_done_in) + (nfull - 1) * B + m
        nblk = (B * sps + 16 + len(chunk)) // (B * sps - 1) + 1      # split_blocks(), carry_cap = 100 sps + 16
        assert nfull + 1 <= nblk
        assert T - P <= B * sps + 16                                   # the carried tail fits carry_cap
        # demod_volume_kernel
        total = nfull * B + m
        va = {}
        for s0 in range(0, total, 32):
            idx = range(s0, min(s0 + 32, total))
            starts = [rec[s // B][0] + (s % B) * sps + (rec[s // B][1] if s % B else 0) for s in idx]
            assert starts == sorted(starts)
            assert starts[-1] + sps <= T and starts[-1] + sps - starts[0] <= 32 * sps + 2
            for s, a in zip(idx, starts):
                va[s] = ("v", self.base + a)
        # demod_slice_kernel
        sym, writers = [None] * emitted, 0
        for b in range(nblk):
            if b > nfull:
                continue
            mb = B if b < nfull else m
            j_done = j_done_in if b == 0 else 0
            if mb <= j_done and nfull != 0:
                continue
            vol = [va[b * B + j] if j < mb else None for j in range(B)]
            pv = st_in["vol_prev"] if b == 0 else [va[(b - 1) * B + j] for j in range(B)]
            for j in range(j_done, mb):
                k = b * B - j_done_in + j
                assert sym[k] is None
                sym[k] = (vol[j], frozenset(vol[:j + 1] + pv[j + 1:]))
            if b == nfull - 1:
                st_out["vol_prev"], writers = vol, writers + 1
            elif nfull == 0:
                st_out["vol_prev"], writers = list(pv), writers + 1
        assert writers == 1 and all(x is not None for x in sym)
        self.state, self.tail = st_out, stream[P:]
        self.base += P
        return sym


def test_split_schedule_equals_one_kernel_walk_on_random_chunkings():
    for seed in range(400):
        rnd = random.Random(seed)
        sps = rnd.choice([5, 10, 12, 20, 40])
        a, b = OneKernel(sps), Split(sps)
        pos = 0
        for it in range(rnd.randint(3, 20)):
            n = rnd.choice([1, 2, 7, sps, sps + 1, sps + 2, 3 * sps, 99 * sps, 100 * sps - 1, 100 * sps, 100 * sps + 1,
                            101 * sps, rnd.randint(1, 500 * sps)])
            chunk = list(range(pos, pos + n))
            pos += n
            assert a.call(chunk) == b.call(chunk), (seed, it, n)
            assert (a.vo, a.j_done, a.vol_prev, a.tail) == (b.state["vo"], b.state["j_done"], b.state["vol_prev"], b.tail)
