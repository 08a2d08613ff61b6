"""Round-2 experiment behind DESIGN.md section 9 ("swap trick"): can the cross-lane stages of the direction-free bitonic
sorter of sort_warp.cuh be done as [exchange half of the registers with the partner lane] + [local compare-exchange],
i.e. 4 shuffles + 8 min/max per lane and stage instead of 8 + 16?

Model: physical index p = lane * EPT + slot; M maps physical -> logical index (linear over GF(2), columns M[j]). A stage
pairs logical i with i ^ mask; with delta = M^-1 mask = (delta_lane, delta_slot) the pair can be brought into one lane by
exchanging the slots whose bit t is set (t in delta_slot) with lane ^ delta_lane. Writing min to the slot with bit t = 0
and max to the other keeps the min/max direction static and turns M into M ^ mask (x) (row_h(M) ^ e_t): still linear, so
the whole network can be planned at compile time. The catch this script demonstrates: that relabelling pins logical bit h
to physical slot bit t, so after three consecutive stages the three slot bits are spoken for and the fourth stage of a
merge pairs elements that differ in lane bits only (delta_slot = 0): no schedule gets past stage 9 of 36, from any initial
layout. At most three stages per merge can be cheap — which the blocked layout already achieves with its register stages.

    python scripts/swap_ce_layout_search.py        # prints the histogram of the stage at which random schedules die
"""
import random


def apply(M, p):
    r, j = 0, 0
    while p:
        if p & 1:
            r ^= M[j]
        p >>= 1
        j += 1
    return r


def inverse_apply(M, y, nb):
    for p in range(1 << nb):
        if apply(M, p) == y:
            return p
    raise ValueError


def stages(nb):
    out, k = [], 2
    while k <= (1 << nb):
        out.append(k - 1)              # mirror stage of the merge of two runs of k/2
        j = k >> 2
        while j > 0:
            out.append(j)              # half-cleaners
            j >>= 1
        k <<= 1
    return out


def plan(EPT, M0, choose):
    SB = EPT.bit_length() - 1
    nb = 5 + SB
    M, sched = list(M0), []
    for si, mask in enumerate(stages(nb)):
        d = inverse_apply(M, mask, nb)
        ds, dl = d & (EPT - 1), d >> SB
        if ds == 0:
            return None, si            # pure-lane pair: needs the classic cross-lane compare-exchange
        t = choose([t for t in range(SB) if (ds >> t) & 1])
        if dl:
            M = list(M)
            M[t] ^= apply(M, dl << SB)
        h = mask.bit_length() - 1
        f = [(M[j] >> h) & 1 for j in range(nb)]
        f[t] ^= 1
        M = [M[j] ^ (mask if f[j] else 0) for j in range(nb)]
        sched.append((dl, t, ds))
    return sched, M


def random_invertible(nb):
    while True:
        M = [random.randrange(1, 1 << nb) for _ in range(nb)]
        rows, rank = list(M), 0
        for bit in range(nb):
            piv = next((i for i in range(rank, nb) if (rows[i] >> bit) & 1), None)
            if piv is None:
                continue
            rows[rank], rows[piv] = rows[piv], rows[rank]
            for i in range(nb):
                if i != rank and (rows[i] >> bit) & 1:
                    rows[i] ^= rows[rank]
            rank += 1
        if rank == nb:
            return M


if __name__ == "__main__":
    random.seed(1)
    EPT, nb = 8, 8
    died, found = {}, 0
    for attempt in range(20000):
        M0 = random_invertible(nb) if attempt % 2 else [1 << j for j in range(nb)]
        sched, res = plan(EPT, M0, random.choice)
        if sched is None:
            died[res] = died.get(res, 0) + 1
        else:
            found += 1
    print("stage masks:", stages(nb))
    print("complete schedules found:", found, "| stage at which the others met a pure-lane pair:", sorted(died.items()))
