"""Host-side model of the record ring of `k_pass1_cond<VT, U, STAGED>` (scdiffcom_b200/csrc/scdc_kernels.cuh).

The kernel walks the cell-type segments of one gene column (`segptr`, pass-1 copy of the matrix: the reference's
`aggregate_cells` order inside every (cell type, condition) group, R/utils_cci.R:441-462) in groups of U records, three
groups in flight, and — STAGED — reads the records from a 64-slot ring that spans the COLUMN: `top_up(base)` makes
[base, base + 32) resident and may overwrite only records below `base`. This model replays exactly that control flow
(same loop bounds, same prefetch distances) with lock-step lanes and checks the two properties the kernel relies on:
every ring read returns the record it asks for, and every record of the column is accumulated exactly once, in order.
No GPU and no oracle involved: it pins the index arithmetic that the GPU parity tests exercise on real shapes."""
import random

import pytest


def replay_column(seg_lens, U, col_beg):
    sp = [col_beg]
    for n in seg_lens:
        sp.append(sp[-1] + n)
    col_end = sp[-1]
    ring = [None] * 64
    filled = col_beg
    ahead = [col_beg + lane if col_beg + lane < col_end else None for lane in range(32)]
    consumed = []

    def top_up(base):
        nonlocal filled, ahead
        while filled < base + 32 and filled < col_end:
            # the half that is overwritten holds records [filled - 64, filled - 32): all below `base`
            assert filled - 32 <= base
            for lane in range(32):
                ring[(filled - col_beg + lane) & 63] = ahead[lane]
            filled += 32
            ahead = [filled + lane if filled + lane < col_end else None for lane in range(32)]

    for t in range(len(seg_lens)):
        beg, end = sp[t], sp[t + 1]
        if end <= beg:
            continue

        def load_recs(e0):
            out = []
            for u in range(U):
                if e0 + u < end:
                    got = ring[(e0 + u - col_beg) & 63]
                    assert got == e0 + u, (got, e0 + u, filled, beg, end)
                    out.append(got)
                else:
                    out.append(None)  # the kernel's zero record: bit-neutral
            return out

        def accumulate(recs):
            consumed.extend(r for r in recs if r is not None)

        top_up(beg)
        r0 = load_recs(beg)
        r1 = load_recs(beg + U)
        base = beg
        while base < end:
            top_up(base)
            r2 = load_recs(base + 2 * U); accumulate(r0)
            r0 = load_recs(base + 3 * U); accumulate(r1)
            r1 = load_recs(base + 4 * U); accumulate(r2)
            base += 3 * U
    assert consumed == list(range(col_beg, col_end))


@pytest.mark.parametrize("U", [2, 4])
def test_ring_serves_every_record_once_and_in_order(U):
    assert 5 * U <= 32  # the kernel's static_assert: one loop iteration reads at most 32 records ahead of its base
    rng = random.Random(20260 + U)
    for _ in range(1500):
        T = rng.randint(1, 12)
        kind = rng.random()
        if kind < 0.3:
            segs = [rng.randint(0, 5) for _ in range(T)]          # many tiny / empty cell types
        elif kind < 0.6:
            segs = [rng.randint(0, 70) for _ in range(T)]
        else:
            segs = [rng.choice([0, 1, 31, 32, 33, 63, 64, 65, 100, 200]) for _ in range(T)]  # ring-size boundaries
        replay_column(segs, U, col_beg=rng.randint(0, 500))


def test_ring_handles_an_empty_column_and_a_single_record():
    replay_column([0, 0, 0], 4, 17)
    replay_column([0, 1, 0], 4, 0)
