"""GPU parity of the whole eval forward (net.py:106-144 through the C ABI) against goldens produced by the
REAL reference and against the oracle: BASELINE.json gates = identical argmax, logits within 1e-5 (fp32) /
1e-3 (bf16) normwise relative."""
import numpy as np
import pytest
import torch

from oracle import qat_oracle as O
from oracle import synth
from oracle.make_golden import FWD_CASES

from gpu_util import TOL, build_model, dev, golden, relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["fwd_b8", "fwd_b3", "fwd_b1", "fwd_k5", "fwd_stress", "fwd_b256", "fwd_t150"])
def test_forward_matches_reference_golden(name, prec):
    g = golden(name)
    B, T, E, K, wseed, bseed = FWD_CASES[name]
    sd = synth.make_state_dict(wseed, num_experts=E)
    model = build_model(sd, E, K, prec)
    for agg in (model.at_aggregator, model.vt_aggregator):
        agg.keep_aux = True
    batch = synth.make_batch(bseed, B, T)
    data = {k: v.to(dev()) for k, v in batch.items()}
    with torch.no_grad():
        out = model(data)["out"]
    ranks_ok = all(agg.last_aux[0].cpu().tolist() == g[nm + ".topk_inds"].astype(np.int64).tolist()
                   for agg, nm in ((model.at_aggregator, "at_aggregator"), (model.vt_aggregator, "vt_aggregator")))
    assert ranks_ok, "router rank lists differ from the reference"
    assert out.shape == (B, 42)
    assert relerr(out, g["out"]) < TOL[prec]
    per_sample = ((out.cpu() - torch.as_tensor(g["out"])).norm(dim=1) / torch.as_tensor(g["out"]).norm(dim=1)).max()
    assert float(per_sample) < 3 * TOL[prec]
    assert out.argmax(1).cpu().tolist() == g["argmax"].tolist()


@pytest.mark.parametrize("name", ["fwd_b8", "fwd_b256"])
def test_forward_with_native_text_tower(name):
    """Full forward (token ids in) with the CLIP tower through qat_clip_text_fwd: same gates as above."""
    g = golden(name)
    B, T, E, K, wseed, bseed = FWD_CASES[name]
    sd = synth.make_state_dict(wseed, num_experts=E)
    model = build_model(sd, E, K, "bf16")
    model.quest_encoder.compute = "native"
    batch = synth.make_batch(bseed, B, T)
    data = {k: v.to(dev()) for k, v in batch.items()}
    with torch.no_grad():
        out = model(data)["out"]
    assert relerr(out, g["out"]) < TOL["bf16"]
    assert out.argmax(1).cpu().tolist() == g["argmax"].tolist()


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_fused_call_matches_block_by_block(prec):
    """qat_forward_fwd (one C-ABI call) and the drop-in modules called one by one run the same arithmetic; in bf16 mode
    the fused path hands planes across block boundaries instead of re-deriving them from fp32 (identical values)."""
    sd = synth.make_state_dict(713, num_experts=7)
    model = build_model(sd, 7, 7, prec)
    batch = synth.make_batch(5, 6, 60)
    data = {k: v.to(dev()) for k, v in batch.items()}
    with torch.no_grad():
        model.fused = True
        fused = model(data)["out"]
        model.fused = False
        blocks = model(data)["out"]
    assert torch.equal(fused, blocks) or relerr(fused, blocks) < 1e-6


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("fused", [True, False])
def test_fp16_patch_feature_store(prec, fused):
    """`patch` may arrive as fp16 (half-size feature store / H2D).  The result must equal the reference
    forward on the same (fp16-representable) values: fp32 mode widens them, bf16 mode feeds them to the tensor cores."""
    sd = synth.make_state_dict(713, num_experts=7)
    model = build_model(sd, 7, 7, prec)
    model.fused = fused
    batch = synth.make_batch(21, 5, 60)
    batch["patch"] = batch["patch"].half()
    ref = O.qa_tiger_forward(sd, {**batch, "patch": batch["patch"].float()})
    data = {k: v.to(dev()) for k, v in batch.items()}
    with torch.no_grad():
        out = model(data)["out"]
    assert relerr(out, ref) < TOL[prec]
    assert out.argmax(1).cpu().tolist() == ref.argmax(1).tolist()


def test_batch_composition_is_part_of_the_function():
    """logits depend on the local batch; DP sharding must therefore be defined on local batches."""
    sd = synth.make_state_dict(713, num_experts=7)
    model = build_model(sd, 7, 7, "fp32")
    batch = synth.make_batch(2, 3, 60)
    data = {k: v.to(dev()) for k, v in batch.items()}
    with torch.no_grad():
        full = model(data)["out"]
        one = model({k: v[:1].contiguous() for k, v in data.items()})["out"]
        ref1 = O.qa_tiger_forward(sd, {k: v[:1] for k, v in batch.items()})
    assert relerr(one, ref1) < 1e-5
    assert relerr(one, full[:1]) > 1e-3


def test_text_tower_matches_oracle():
    sd = synth.make_state_dict(713, num_experts=7)
    model = build_model(sd, 7, 7, "fp32")
    batch = synth.make_batch(9, 6, 60)
    model.quest_encoder.compute = "fp32"          # torch fp32 matmuls (the reference's arithmetic)
    with torch.no_grad():
        q, w = model.quest_encoder(batch["quest"].to(dev()))
        rq, rw = O.clip_text_encoder(sd, "quest_encoder", batch["quest"])
    assert relerr(q, rq) < 2e-5 and relerr(w, rw) < 2e-5
    model.quest_encoder.compute = "bf16x3"        # Linear layers through the tcgen05 bf16 hi/lo GEMM
    with torch.no_grad():
        q3, w3 = model.quest_encoder(batch["quest"].to(dev()))
    assert relerr(q3, rq) < 1e-4 and relerr(w3, rw) < 1e-4
    # the whole tower in one C-ABI call (qat_clip_text_fwd): fp32 FFMA and bf16 hi+lo x3 tensor-core modes
    for nprec, tol in (("fp32", 2e-5), ("bf16", 1e-4)):
        model.quest_encoder.compute = "native"
        model.quest_encoder.native_precision = nprec
        with torch.no_grad():
            qn, wn = model.quest_encoder(batch["quest"].to(dev()))
        assert qn.shape == q.shape and wn.shape == w.shape
        assert relerr(qn, rq) < tol and relerr(wn, rw) < tol, (nprec, relerr(qn, rq), relerr(wn, rw))
    model.quest_encoder.compute = "fp32"
    model.quest_encoder.dedup = True
    rep = batch["quest"][[0, 1, 0, 2, 1, 0]].to(dev())
    with torch.no_grad():
        q2, w2 = model.quest_encoder(rep)
    assert relerr(q2[2], q[0]) < 1e-5 and relerr(w2[4], w[1]) < 1e-5
    # cross-call memo: second call sees only cached rows (and one new question), results identical to the direct path
    model.quest_encoder.memo = True
    with torch.no_grad():
        q3a, w3a = model.quest_encoder(rep)
        q3b, w3b = model.quest_encoder(torch.cat([rep[:3], batch["quest"][3:4].to(dev())]))
    assert relerr(q3a, q2) < 1e-6 and relerr(w3a, w2) < 1e-6
    assert relerr(q3b[:3], q2[:3]) < 1e-6 and relerr(q3b[3], q[3]) < 1e-5 and relerr(w3b[3], w[3]) < 1e-5
    assert len(model.quest_encoder._memo_state["slots"]) == 4
    model.quest_encoder.memo = False


def test_state_dict_is_reference_compatible():
    sd = synth.make_state_dict(713, num_experts=7)
    model = build_model(sd, 7, 7)
    keys = set(model.state_dict().keys())
    assert keys == set(sd.keys()) and len(keys) == 286
    for k, v in model.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_stress_config_at_its_own_batch(prec):
    """BASELINE configs[4] at its per-GPU batch: B = 128, T = 120, E = K = 14 against the reference golden (logits + ranks)."""
    g = golden("fwd_stress_b128")
    B, T, E, K, wseed, bseed = FWD_CASES["fwd_stress_b128"]
    sd = synth.make_state_dict(wseed, num_experts=E)
    model = build_model(sd, E, K, prec)
    for agg in (model.at_aggregator, model.vt_aggregator):
        agg.keep_aux = True
    batch = synth.make_batch(bseed, B, T)
    data = {k: v.to(dev()) for k, v in batch.items()}
    with torch.no_grad():
        out = model(data)["out"]
    for agg, nm in ((model.at_aggregator, "at_aggregator"), (model.vt_aggregator, "vt_aggregator")):
        assert agg.last_aux[0].cpu().tolist() == g[nm + ".topk_inds"].astype(np.int64).tolist()
    assert relerr(out, g["out"]) < TOL[prec]
    assert out.argmax(1).cpu().tolist() == g["argmax"].tolist()


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("B,T", [(3, 150), (2, 40)])
def test_forward_any_frame_count_matches_oracle(prec, B, T):
    """net.py:106-144 is length-agnostic in T (modules.py:217-221, 562-578): frame counts past the 128-token attention cores
    (and short ones off the fused expert epilogue's tile) are compared with the oracle on the same batch."""
    sd = synth.make_state_dict(713, num_experts=7)
    model = build_model(sd, 7, 7, prec)
    batch = synth.make_batch(17 + T, B, T)
    with torch.no_grad():
        ref = O.qa_tiger_forward(sd, batch, K=7, E=7)
        out = model({k: v.to(dev()) for k, v in batch.items()})["out"]
    assert relerr(out, ref) < TOL[prec]
    assert out.argmax(1).cpu().tolist() == ref.argmax(1).tolist()


def test_two_threads_two_streams_are_reentrant():
    """nn.DataParallel (trainutils.py:107-108) drives one Python thread per replica through the same library: two threads,
    each with its own stream, model replica and batch, run concurrently and must reproduce their single-threaded logits."""
    import threading
    sd = synth.make_state_dict(713, num_experts=7)
    models = [build_model(sd, 7, 7, "bf16") for _ in range(2)]
    batches = [{k: v.to(dev()) for k, v in synth.make_batch(40 + i, 5 + i, 60).items()} for i in range(2)]
    with torch.no_grad():
        want = [m(b)["out"].clone() for m, b in zip(models, batches)]
    torch.cuda.synchronize()
    got, errs = [None, None], []

    def work(i):
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s), torch.no_grad():
                for _ in range(6):
                    o = models[i](batches[i])["out"]
                s.synchronize()
            got[i] = o
        except Exception as e:   # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for i in range(2):
        assert torch.equal(got[i], want[i])


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_nn_dataparallel_wrapper(prec):
    """trainutils.py:107-108 wraps the model in nn.DataParallel when it is not distributed: replicate() + one host thread per
    replica.  The logits depend on the LOCAL batch, so the wrapped forward equals the plain model run on each
    scattered chunk.  On a one-GPU box both replicas sit on device 0 (two host threads on one stream); with two or more
    GPUs the same check also runs across devices 0 and 1."""
    sd = synth.make_state_dict(713, num_experts=7)
    model = build_model(sd, 7, 7, prec)
    batch = synth.make_batch(23, 8, 60)
    data = {k: v.to(dev()) for k, v in batch.items()}
    with torch.no_grad():
        halves = [model({k: v[i * 4:(i + 1) * 4].contiguous() for k, v in data.items()})["out"] for i in range(2)]
    want = torch.cat(halves)
    id_sets = [[0, 0]] + ([[0, 1]] if torch.cuda.device_count() >= 2 else [])
    for ids in id_sets:
        dp = torch.nn.DataParallel(model, device_ids=ids, output_device=0)
        with torch.no_grad():
            got = dp(data)["out"]
        torch.cuda.synchronize()
        assert got.shape == (8, 42)
        assert torch.equal(got, want), f"DataParallel over devices {ids} differs from the per-chunk forwards"


def test_data_writes_need_invalidate_prepared():
    """ADVICE: derived weights are cached on (data_ptr, _version); load_state_dict / .to() invalidate them automatically,
    writes through `.data` need qa_tiger_b200._cabi.invalidate_prepared()."""
    from qa_tiger_b200 import _cabi
    sd = synth.make_state_dict(713, num_experts=7)
    model = build_model(sd, 7, 7, "bf16")
    data = {k: v.to(dev()) for k, v in synth.make_batch(8, 4, 60).items()}
    with torch.no_grad():
        a = model(data)["out"].clone()
        sd2 = {k: (v * 1.01 if k.startswith("crs_attn.linear1") else v) for k, v in sd.items()}
        model.load_state_dict(sd2)                       # same storage, no _version bump on the module side needed
        b = model(data)["out"].clone()
        assert relerr(b, a) > 1e-5
        model.crs_attn.linear1.weight.data.copy_(sd["crs_attn.linear1.weight"])
        model.crs_attn.linear1.bias.data.copy_(sd["crs_attn.linear1.bias"])
        _cabi.invalidate_prepared()
        c = model(data)["out"]
    assert torch.equal(c, a)
