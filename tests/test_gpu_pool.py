"""GPU: wb_pool — one handle, several batches in flight, one host thread.  Every batch must come back exactly as a
context running alone returns it (same tokens, same file indices), whatever the order the tickets are collected in;
a failing batch reports its own error text on the collecting thread and leaves the pool usable."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
PROMPT = [50258, 50259, 50359, 50363]
EOT = 50257


def test_pool_matches_solo_context_and_tickets_collect_in_any_order(wb):
    B, n_new = 8, 20
    batches = [wb.synth.batch(B, seed=50 + i, seconds=30.0) for i in range(5)]
    cfg = wb.default_cfg("base", precision=wb.WB_PREC_BF16, max_batch=B, max_chunks=B)
    solo = wb.Whisper(cfg)
    want = [solo.transcribe_batch(list(b), PROMPT, n_new, EOT) for b in batches]
    solo.close()
    pool = wb.Pool(wb.default_cfg("base", precision=wb.WB_PREC_BF16, max_batch=B, max_chunks=B), 3)
    assert pool.slots == 3
    for rnd in range(2):
        tickets = [pool.submit(list(b), PROMPT, n_new, EOT) for b in batches]      # 5 batches through 3 slots, one thread
        assert len(set(tickets)) == 5
        order = [3, 0, 4, 1, 2] if rnd else [0, 1, 2, 3, 4]
        for k in order:
            toks, fidx = pool.wait(tickets[k])
            assert toks == want[k][0]
            assert fidx.tolist() == want[k][1].tolist()
    with pytest.raises(wb.WbError, match="unknown ticket"):
        pool.wait(tickets[0])                                                       # already collected
    pool.close()


def test_pool_reports_a_failed_batch_and_keeps_working(wb):
    B = 4
    pool = wb.Pool(wb.default_cfg("toy", max_batch=B, max_chunks=B), 2)
    good = wb.synth.batch(B, seed=7, seconds=6.0)
    too_many = wb.synth.batch(B + 1, seed=8, seconds=6.0)                           # 5 chunks > max_chunks 4
    t_bad = pool.submit(list(too_many), [1, 2, 3, 4], 5, 1030)
    t_ok = pool.submit(list(good), [1, 2, 3, 4], 5, 1030)
    with pytest.raises(wb.WbError, match="exceed max_chunks"):
        pool.wait(t_bad)
    toks, fidx = pool.wait(t_ok)
    assert len(toks) == B and fidx.tolist() == list(range(B))
    solo = wb.Whisper(wb.default_cfg("toy", max_batch=B, max_chunks=B))
    assert solo.transcribe_batch(list(good), [1, 2, 3, 4], 5, 1030)[0] == toks
    solo.close()
    pool.close()
