"""CPU: an executable model of the work decomposition of csr_stream_kernel (hpr-lp-c_b200/csrc/kernels.cuh) -- the index
logic only, restated line by line in Python with small item sizes so that every boundary case occurs on small matrices:

  * item_row (build_item_rows_kernel), the rows an item touches, head / cont classification;
  * which partial sums are published, to the CTA-local (shared-memory) or to the global packet, and which ones the item
    that holds the END of a cut row consumes (look-back) -- every packet must be published exactly once before it is
    consumed, consumed exactly once, and every row completed exactly once with the right sum;
  * column bands (Engine::build_bands): carry-in / carry-out over the bands gives the same row sums;
  * CTA-range launches (CsrView::cta0 / n_launch): the property the pipelined exchange of the partitioned x-phase relies on
    -- after the launches up to range c, every row below the common column boundary is complete.

The GPU tests check the kernels themselves; this model pins the invariants the kernel's control flow is built on."""
import numpy as np
import pytest


def item_rows(row_ptr, n_witems, chunk):
    """build_item_rows_kernel: item_row[i] = first r with rowPtr[r+1] > i*chunk; 0 for i = 0; `rows` at or beyond nnz."""
    rows, nnz = len(row_ptr) - 1, int(row_ptr[-1])
    out = np.zeros(n_witems + 1, np.int64)
    for i in range(1, n_witems + 1):
        target = i * chunk
        out[i] = rows if target >= nnz else int(np.searchsorted(row_ptr[1:], target, side="right"))
    return out


def run_pass(row_ptr, prod, warp_chunk, warps, carry_in=None, want_carry_out=False, cta_range=None, state=None):
    """One launch of the model kernel over CTAs cta_range (default: all).  prod[k] = product of nonzero k.
    Returns (completed: dict row -> sum, state) where state holds the global packets between launches."""
    rows, nnz = len(row_ptr) - 1, int(row_ptr[-1])
    chunk = warp_chunk * warps                                   # kChunk
    n_ctas = max(1, -(-nnz // chunk))                            # n_items
    n_witems = n_ctas * warps
    irow = item_rows(row_ptr, n_witems, warp_chunk)
    if state is None:
        state = dict(head={}, tail={})                           # global packets: item -> value (absent = "not published")
    lo_cta, hi_cta = cta_range if cta_range is not None else (0, n_ctas)
    completed = {}

    def complete(r, total):
        assert r not in completed, f"row {r} completed twice"
        if carry_in is not None:
            total = total + carry_in[r]
        completed[r] = total

    for cta in range(lo_cta, hi_cta):
        cta_part = {}                                            # (warp, 'head'|'tail') -> value : shared-memory hand-off
        cta_item0, cta_end = cta * warps, (cta + 1) * chunk
        finishers = []
        # pass A: every warp's phase 2 (publishes come before any wait, so the order of the warps does not matter)
        for warp in np.random.default_rng(cta).permutation(warps):
            item = cta * warps + int(warp)
            s = item * warp_chunk
            e = min(s + warp_chunk, nnz) if s < nnz else s
            rA, rB = int(irow[item]), int(irow[item + 1])
            r_last = rB if rB < rows else rows - 1
            for r in range(rA, r_last + 1):
                p0, p1 = int(row_ptr[r]), int(row_ptr[r + 1])
                a, b = max(p0, s), min(p1, e)
                tot = float(prod[a:b].sum()) if b > a else 0.0
                head = (r == rA) and (p0 < s)
                cont = p1 > e
                if not head and not cont:
                    complete(r, tot)
                elif head and not cont:
                    finishers.append((int(warp), item, r, p0, tot))
                elif head or p0 < e:
                    kind = "head" if head else "tail"
                    if p1 <= cta_end:
                        assert (int(warp), kind) not in cta_part
                        cta_part[(int(warp), kind)] = tot
                    else:
                        assert item not in state[kind], "global packet published twice"
                        state[kind][item] = tot
        # pass B: the rows that end in an item after entering it from the left
        for warp, item, r, p0, own in finishers:
            ia, ib = p0 // warp_chunk, item
            assert ia < ib
            total = None
            for j in range(ia, ib):
                kind = "tail" if j == ia else "head"
                if j >= cta_item0:
                    v = cta_part.pop((j - cta_item0, kind))      # KeyError = waiting for a packet nobody publishes
                else:
                    v = state[kind].pop(j)
                total = v if total is None else total + v
            complete(r, total + own)
        assert not cta_part, f"unconsumed shared packets {cta_part}"
    return completed, state


def random_row_ptr(rng, rows, kind):
    if kind == "short":
        lens = rng.integers(0, 6, rows)
    elif kind == "long":
        lens = rng.integers(0, 3, rows)
        for r in rng.choice(rows, max(1, rows // 10), replace=False):
            lens[r] = rng.integers(20, 200)
    elif kind == "empty_edges":
        lens = rng.integers(0, 12, rows)
        lens[:3] = 0
        lens[-4:] = 0
        lens[rng.choice(rows, rows // 3, replace=False)] = 0
    else:  # exact multiples: rows ending exactly on item boundaries
        lens = rng.choice([0, 4, 8, 16], rows)
    rp = np.zeros(rows + 1, np.int64)
    rp[1:] = np.cumsum(lens)
    if rp[-1] == 0:
        rp[-1:] = 1
        rp[1:] = np.maximum(rp[1:], 1)   # at least one nonzero
        rp[1:] = 1
    return rp


@pytest.mark.parametrize("kind", ["short", "long", "empty_edges", "multiples"])
@pytest.mark.parametrize("warp_chunk,warps", [(8, 2), (4, 4), (16, 1)])
def test_every_row_completed_once_with_the_right_sum(kind, warp_chunk, warps):
    for seed in range(12):
        rng = np.random.default_rng(1000 * seed + warp_chunk)
        rows = int(rng.integers(1, 120))
        rp = random_row_ptr(rng, rows, kind)
        nnz = int(rp[-1])
        prod = rng.integers(-9, 10, nnz).astype(float)            # integers: sums are exact in any order
        done, state = run_pass(rp, prod, warp_chunk, warps)
        assert sorted(done) == list(range(rows)), (kind, seed, sorted(set(range(rows)) - set(done)))
        want = np.array([prod[rp[r]:rp[r + 1]].sum() for r in range(rows)])
        assert np.array_equal(np.array([done[r] for r in range(rows)]), want)
        assert not state["head"] and not state["tail"], "global packets left behind (the next launch would misread them)"


@pytest.mark.parametrize("n_bands", [2, 5, 40])
def test_column_bands_carry_gives_the_same_row_sums(n_bands):
    rng = np.random.default_rng(7)
    rows, cols, warp_chunk, warps = 60, 40, 8, 2
    lens = rng.integers(0, 15, rows)
    rp = np.zeros(rows + 1, np.int64); rp[1:] = np.cumsum(lens)
    col = np.concatenate([np.sort(rng.choice(cols, int(k), replace=False)) for k in lens]).astype(np.int64)
    prod = rng.integers(-9, 10, int(rp[-1])).astype(float)
    want = np.array([prod[rp[r]:rp[r + 1]].sum() for r in range(rows)])
    band_cols = -(-cols // n_bands)
    carry = None
    for b in range(n_bands):                                      # band b: entries with col // band_cols == b, order kept
        keep = (col // band_cols) == b
        brp = np.zeros(rows + 1, np.int64)
        brp[1:] = np.cumsum([int(keep[rp[r]:rp[r + 1]].sum()) for r in range(rows)])
        bprod = prod[keep]
        # (an empty band -- no entry in that column slice -- still has to pass every row's carry on: one item owns all rows)
        done, state = run_pass(brp, bprod, warp_chunk, warps, carry_in=carry)
        assert not state["head"] and not state["tail"]
        assert sorted(done) == list(range(rows))
        carry = np.array([done[r] for r in range(rows)])
    assert np.array_equal(carry, want)


@pytest.mark.parametrize("n_ranges", [2, 3, 5])
def test_cta_range_launches_complete_the_common_column_ranges(n_ranges):
    """Pipelined exchange (git branch pipelined-exchange): rows [n c / C, n (c+1) / C) must be complete once the CTAs up to
    ceil(rowPtr[n (c+1) / C] / kChunk) have run -- on every GPU, whatever its own nonzero distribution."""
    warp_chunk, warps = 8, 2
    chunk = warp_chunk * warps
    for seed in range(10):
        rng = np.random.default_rng(50 + seed)
        rows = int(rng.integers(n_ranges * 4, 150))
        rp = random_row_ptr(rng, rows, ["short", "long", "empty_edges", "multiples"][seed % 4])
        nnz = int(rp[-1])
        prod = rng.integers(-9, 10, nnz).astype(float)
        n_ctas = max(1, -(-nnz // chunk))
        bound_row = [rows * c // n_ranges for c in range(n_ranges + 1)]
        bound_cta = [min(n_ctas, -(-int(rp[r]) // chunk)) for r in bound_row]
        bound_cta[0], bound_cta[-1] = 0, n_ctas
        for c in range(1, n_ranges + 1):
            bound_cta[c] = max(bound_cta[c], bound_cta[c - 1])
        done_all, state = {}, None
        for c in range(n_ranges):
            done, state = run_pass(rp, prod, warp_chunk, warps, cta_range=(bound_cta[c], bound_cta[c + 1]), state=state)
            assert not (set(done) & set(done_all))
            done_all.update(done)
            missing = [r for r in range(bound_row[c + 1]) if r not in done_all]
            assert not missing, (seed, c, missing[:5])
        want = np.array([prod[rp[r]:rp[r + 1]].sum() for r in range(rows)])
        assert np.array_equal(np.array([done_all[r] for r in range(rows)]), want)
        assert not state["head"] and not state["tail"]


def test_model_is_sensitive_to_the_boundary_rule():
    """Control: with floor instead of ceil for the CTA boundary, rows ending inside the boundary CTA are not complete when the
    range's collective would start -- the model must notice (this is what makes the test above meaningful)."""
    warp_chunk, warps, n_ranges = 8, 2, 3
    chunk = warp_chunk * warps
    caught = 0
    for seed in range(20):
        rng = np.random.default_rng(500 + seed)
        rows = int(rng.integers(20, 150))
        rp = random_row_ptr(rng, rows, "short")
        nnz = int(rp[-1])
        prod = np.ones(nnz)
        n_ctas = max(1, -(-nnz // chunk))
        bound_row = [rows * c // n_ranges for c in range(n_ranges + 1)]
        bound_cta = [min(n_ctas, int(rp[r]) // chunk) for r in bound_row]      # floor: wrong
        bound_cta[0], bound_cta[-1] = 0, n_ctas
        done_all, state = {}, None
        for c in range(n_ranges):
            done, state = run_pass(rp, prod, warp_chunk, warps, cta_range=(bound_cta[c], max(bound_cta[c + 1], bound_cta[c])), state=state)
            done_all.update(done)
            if any(r not in done_all for r in range(bound_row[c + 1])):
                caught += 1
                break
    assert caught > 0
