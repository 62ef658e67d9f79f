"""N > 1 on real GPUs (skipped on single-GPU boxes; the host-side logic is covered by tests/test_distributed_cpu.py with gloo):
  * the native multi-GPU context of the C ABI (lsqr_ctx_create_multi: one process, NCCL inside the library) against a
    single-GPU context: identical counts, winner, consensus set and compute() result, refit within the summation-order tolerance;
  * tools/check_multi_gpu.py under torchrun, one process per GPU, with the library's own communicator (lsqr_ctx_init_nccl)
    and with caller-supplied hooks (lsqr_set_shard + torch.distributed): sharded requests identical to unsharded ones."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_ranks_match_one():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29571", os.path.join(ROOT, "tools", "check_multi_gpu.py")], capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTI-GPU CHECK OK" in out.stdout, out.stdout[-3000:]


def test_native_multi_gpu_context_matches_single_gpu():
    """SURVEY.md 8b ctx_create(ngpus): every C-ABI entry point on a context spanning all GPUs returns what one GPU returns."""
    import numpy as np
    import torch
    from lsqrrecipes_b200 import FP32, FP64, Engine, synth
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least two GPUs")
    for name, n, H in (("plane3", 1_500_003, 300_000), ("sphere3", 400_001, 50_000), ("absor", 200_001, 20_000), ("pivot", 50_000, 4_000)):
        data, _ = synth.GENERATORS[name](n, seed=3)
        delta = synth.DELTAS[name]
        one, many = Engine(name, delta), Engine(name, delta, gpus=0)
        assert many.world == ngpu and one.world == 1
        one.upload(data)
        many.upload(data)
        for precision, h in ((FP32, H), (FP64, min(H, 4096))):
            a = one.score(count=h, seed=11, precision=precision, want_counts=True, want_params=True)
            b = many.score(count=h, seed=11, precision=precision, want_counts=True, want_params=True)
            assert np.array_equal(a["counts"], b["counts"]) and np.array_equal(np.nan_to_num(a["params"]), np.nan_to_num(b["params"]))
            assert (a["best_index"], a["best_count"], a["n_valid"]) == (b["best_index"], b["best_count"], b["n_valid"])
            assert np.array_equal(a["best_params"], b["best_params"]) and np.array_equal(a["best_subset"], b["best_subset"])
        assert one.consensus(a["best_params"]) == many.consensus(a["best_params"])
        assert np.array_equal(one.get_mask(), many.get_mask()) and np.array_equal(one.get_mask_bits(), many.get_mask_bits())
        assert np.array_equal(one.get_mask(), one.get_mask_bits())
        assert np.allclose(one.refine(), many.refine(), rtol=1e-9, atol=1e-9)
        ra, rb = one.ransac(0.999, precision=FP32, seed=12), many.ransac(0.999, precision=FP32, seed=12)
        assert (ra["best_index"], ra["fraction"], ra["tries"]) == (rb["best_index"], rb["fraction"], rb["tries"])
        assert np.array_equal(ra["mask"], rb["mask"]) and np.allclose(ra["params"], rb["params"], rtol=1e-9, atol=1e-9)
        # lsqr_compute: the pipelined call with the data on the host (interleaved chunk upload + all-gather inside the library)
        ca, cb = one.compute(data, 0.999, precision=FP32, seed=12), many.compute(data, 0.999, precision=FP32, seed=12)
        assert (ca["best_index"], ca["fraction"], ca["tries"]) == (ra["best_index"], ra["fraction"], ra["tries"]) == (cb["best_index"], cb["fraction"], cb["tries"])
        assert np.array_equal(ca["mask"], ra["mask"]) and np.array_equal(cb["mask"], ra["mask"])
        assert np.array_equal(ca["params"], ra["params"]) and np.allclose(cb["params"], ra["params"], rtol=1e-9, atol=1e-9)
        one.close()
        many.close()
    # batched small problems: partitioned over the GPUs, no collective, same answers
    nprob, npts = 4096, 128
    base = [synth.GENERATORS["line2d"](npts, seed=300 + i)[0] for i in range(64)]
    data = np.concatenate([base[i % 64] for i in range(nprob)])
    offsets = (np.arange(nprob + 1) * npts).astype(np.uint64)
    one, many = Engine("line2d", 0.5), Engine("line2d", 0.5, gpus=0)
    a = one.ransac_batch(data, offsets, prob=0.999, max_tries=1024, seed=7, want_masks=True)
    b = many.ransac_batch(data, offsets, prob=0.999, max_tries=1024, seed=7, want_masks=True)
    assert np.array_equal(a["counts"], b["counts"]) and np.array_equal(a["masks"], b["masks"])
    assert np.array_equal(np.nan_to_num(a["params"]), np.nan_to_num(b["params"]))
    one.close()
    many.close()
