"""Multi-GPU tests (need >= 2 GPUs; skipped on a one-GPU box): the node's GPUs behind one C handle (friq_multi_*: contiguous agent
blocks, three grouped ncclAllGather calls, NCCL loaded at run time) and the torchrun / torch.distributed NCCL path bench.py uses.
No gloo anywhere: the collectives run over NVLink between real devices."""
import os
import subprocess
import sys

import numpy as np
import pytest

import _oracle

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fq():
    import friq_b200
    friq_b200.lib()
    return friq_b200


def test_multi_handle_on_one_gpu_equals_plain_batch(fq):
    """friq_multi with ONE device (no collective): same agents, same results as a plain batch; statistics kernel checked."""
    n = 300
    m = fq.Multi("mountaincar", n, 192, n_devices=1, no_random=0)
    P = fq.Problem("mountaincar", no_random=0)
    b = fq.Batch(P, n, 192)
    for _ in range(3):
        m.run(200, fq.RUN_LEARN | fq.RUN_RETILE); b.run(200, fq.RUN_LEARN | fq.RUN_RETILE)
    g = m.gather(); o = b.gather()
    assert np.array_equal(g["rules"], o["rules"]) and np.array_equal(g["last_return"], o["last_return"]) and np.array_equal(g["last_steps"], o["last_steps"])
    assert g["stats"][0] == o["rules"].sum() and g["stats"][2] == o["rules"].max() and np.isclose(g["stats"][1], o["last_return"].sum(), rtol=1e-12)
    m.close(); b.close(); P.close()


def test_multi_gpu_native_handle_matches_single_batch_and_oracle(fq):
    """All GPUs of the box through friq_multi_*: the sharded run equals ONE big batch agent for agent (Philox key = global agent
    index), the NCCL-gathered arrays are in global agent order, a sampled agent of every GPU equals the oracle, and a
    cross-GPU merge_rb (friq_multi_merge_from) equals the oracle's merge."""
    G = fq.device_count()
    if G < 2:
        pytest.skip("needs >= 2 GPUs")
    n, steps = 1000 * G + 37, 600                      # uneven blocks: the last GPU's block is shorter (padding in the gather)
    m = fq.Multi("mountaincar", n, 256, no_random=0)
    assert m.G == G
    eps = np.random.default_rng(5).choice([0.0, 0.01, 0.05], n)
    m.set_hyper(np.full(n, 0.5), np.ones(n), eps)
    P = fq.Problem("mountaincar", no_random=0)
    b = fq.Batch(P, n, 256); b.set_hyper(np.full(n, 0.5), np.ones(n), eps)
    for _ in range(3):
        m.run(steps // 3, fq.RUN_LEARN | fq.RUN_KEEP_WEIGHTS); b.run(steps // 3, fq.RUN_LEARN | fq.RUN_KEEP_WEIGHTS)
    g = m.gather(); o = b.gather()
    for k in ("rules", "last_return", "last_steps"):
        assert np.array_equal(g[k], o[k]), k
    assert g["stats"][0] == o["rules"].sum() and g["stats"][2] == o["rules"].max()
    OP = _oracle.Problem(0, trig=1); OP.set_no_random(0)
    picks = [m.first_agent(i) + 3 for i in range(G)] + [n - 1]
    agents = {}
    for i in picks:
        a = _oracle.Agent(OP, 256, seed=0x5EED0000 + i); a.set_hyper(0.5, 1.0, eps[i]); a.run_steps(steps)
        agents[i] = a
        gpu = max(j for j in range(G) if m.first_agent(j) <= i)
        R, rant, rconc = m.batch(gpu, P).get_rulebase(i - m.first_agent(gpu))
        oR, orant, orconc = a.rb()
        assert R == oR and np.array_equal(rant, orant) and np.array_equal(rconc, orconc), i
    # knowledge sharing across GPUs: agent (first agent of the LAST GPU + 3)'s rules into agent 3 on GPU 0
    src, dst = picks[G - 1], picks[0]
    m.merge_from(src, dst)
    sR, srant, srconc = agents[src].rb()
    agents[dst].merge_rb(srant, srconc)
    R, rant, rconc = m.batch(0, P).get_rulebase(dst)
    oR, orant, orconc = agents[dst].rb()
    assert R == oR and np.array_equal(rant, orant) and np.array_equal(rconc, orconc)
    m.close(); b.close(); P.close()


def test_torchrun_nccl_result_exchange_and_rulebase_gather(tmp_path):
    """bench.py's multi-process path on real GPUs: 2 ranks under torchrun, friq_dist.ResultExchange over NCCL (side stream,
    double-buffered) and friq_dist.gather_rulebases + friq_batch_merge_rb across the ranks, checked against one big batch and the
    oracle inside the worker (tests/_nccl_worker.py)."""
    import friq_b200
    if friq_b200.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(29600 + os.getpid() % 300), os.path.join(REPO, "tests", "_nccl_worker.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "NCCL_WORKER_OK rank 0" in r.stdout and "NCCL_WORKER_OK rank 1" in r.stdout
