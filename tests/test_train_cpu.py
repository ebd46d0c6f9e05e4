"""CPU checks of the fusion-head training step: the oracle (oracle/train_oracle.py) against the fixture generated
from the UNMODIFIED reference (tests/golden/train_head.npz, oracle/make_golden_train.py), and the host logic of the
step (flat buffers, gradient all-reduce on 2 gloo ranks, configuration checks, loud failure without CUDA)."""
import json
import os
import subprocess
import sys

import pytest
import torch

import b200path
import golden_util as gu
from oracle import params as op
from oracle import train_oracle as to

ROOT = b200path.ROOT


def _setup(name="train_head.npz"):
    gold = gu.load(name)
    hp = json.loads(str(gold["hp"]))
    import parameters_default as pd

    params = pd.default_parameters()
    shapes = gu.load_shapes("cnn")["fusion"]
    sd = op.seeded_state_dict(shapes, seed=hp["weight_seed"])
    batch = op.synthetic_head_batch(hp["n"], seed=hp["seed"])
    return gold, hp, params, sd, batch


def check_updated_parameter(gold, name, got, rtol):
    """Parameters after the AdamW steps.  The key bias of nn.MultiheadAttention has a mathematically zero gradient
    (softmax is invariant to a per-query shift of the scores); Adam normalises its rounding noise (1e-10) to full
    +-lr steps, so that third of in_proj_bias is a random walk in the reference too and is not compared."""
    if name.endswith("cross_attn.in_proj_bias"):
        ref = torch.from_numpy(gold[f"param/{name}/full"])
        c = ref.numel() // 3
        keep = torch.cat([torch.arange(0, c), torch.arange(2 * c, 3 * c)])
        got = got.detach().float().cpu()
        err = (got[keep] - ref[keep]).abs().max().item() / ref[keep].abs().max().item()
        assert err <= rtol, f"{name} (query / value thirds): {err:.3e} > {rtol:.1e}"
        return err
    return gu.check(gold, f"param/{name}", got, rtol, what="updated parameter ")


def test_oracle_loss_and_gradients_match_the_reference():
    gold, hp, params, sd, batch = _setup()
    loss, logits, grads = to.head_loss_and_grads(sd, params, *batch, hp["smoothing"], hp["gamma"],
                                                 torch.tensor(hp["class_weights"]))
    assert abs(float(loss) - gold["losses"][0]) <= 1e-5 * abs(gold["losses"][0])
    gu.check(gold, "logits", logits, 1e-5)
    names = sorted(k[len("grad/"):-len("/meta")] for k in gold.files if k.startswith("grad/") and k.endswith("/meta"))
    assert names == sorted(grads) == sorted(hp["updated"])  # the same 20 tensors receive a gradient
    for k in names:
        gu.check(gold, f"grad/{k}", grads[k], 2e-4, what="gradient ")


def test_oracle_adamw_steps_match_the_reference():
    gold, hp, params, sd, batch = _setup()
    losses, new_sd, names = to.train_steps(sd, params, batch, hp["steps"], hp["smoothing"], hp["gamma"],
                                           torch.tensor(hp["class_weights"]), hp["lr"], tuple(hp["betas"]),
                                           hp["eps"], hp["weight_decay"])
    assert sorted(names) == sorted(hp["updated"])
    for a, b in zip(losses, gold["losses"]):
        assert abs(a - b) <= 2e-4 * abs(b), (losses, gold["losses"])
    assert losses[-1] < losses[0]
    for k in names:
        check_updated_parameter(gold, k, new_sd[k], 1e-4)
    untouched = [k for k in sd if k not in names and sd[k].is_floating_point()]
    assert untouched and all(torch.equal(sd[k], new_sd[k]) for k in untouched)


@pytest.mark.parametrize("fixture", ["train_head_mask.npz", "train_head_mask_bce.npz"])
def test_oracle_mask_term_matches_the_reference(fixture):
    """classification + lambda_mask * mean of three mask terms (train_fusion.py:238-255; SoftDiceLoss or DiceBCELoss,
    selector_helpers.py:95-114): 24 tensors get a gradient."""
    gold, hp, params, sd, batch = _setup(fixture)
    masks = op.synthetic_raw(hp["n"], seed=hp["seed"] + 1, kind="S")[2]
    cw = torch.tensor(hp["class_weights"])
    loss, logits, grads = to.head_loss_and_grads(sd, params, *batch, hp["smoothing"], hp["gamma"], cw, masks,
                                                 hp["lambda_mask"], hp["mask_loss_type"])
    assert abs(float(loss) - gold["losses"][0]) <= 1e-5 * abs(gold["losses"][0])
    assert sorted(grads) == sorted(hp["updated"]) and len(grads) == 24
    for k in grads:
        gu.check(gold, f"grad/{k}", grads[k], 2e-4, what="gradient ")
    if hp["mask_loss_type"] != "dice":
        return  # (the optimisation trajectory is checked once, on the default mask loss)
    losses, new_sd, names = to.train_steps(sd, params, batch, hp["steps"], hp["smoothing"], hp["gamma"], cw, hp["lr"],
                                           tuple(hp["betas"]), hp["eps"], hp["weight_decay"], masks, hp["lambda_mask"],
                                           hp["mask_loss_type"])
    for a, b in zip(losses, gold["losses"]):
        assert abs(a - b) <= 2e-4 * abs(b)
    for k in names:
        check_updated_parameter(gold, k, new_sd[k], 1e-4)


def test_flat_views_are_aligned_views():
    from fusion_train import ALIGN, flat_size, flat_views

    ts = [torch.zeros(2, 3), torch.zeros(70), torch.zeros(1, 1, 4)]
    assert flat_size(ts) == 4 * ALIGN and flat_size(ts, align=1) == 80
    flat = torch.arange(4 * ALIGN, dtype=torch.float32)
    v = flat_views(ts, flat)
    assert [tuple(x.shape) for x in v] == [(2, 3), (70,), (1, 1, 4)]
    assert v[1][0] == ALIGN and v[2].flatten()[0] == 3 * ALIGN      # every tensor starts on a 256-byte boundary
    assert all(x.data_ptr() % (4 * ALIGN) == flat.data_ptr() % (4 * ALIGN) for x in v)
    v[0].zero_()
    assert flat[:6].abs().sum() == 0  # views, not copies
    packed = flat_views(ts, torch.arange(80, dtype=torch.float32), align=1)
    assert packed[1][0] == 6 and packed[2].flatten()[0] == 76


def test_split_k_fills_the_machine_without_empty_slices():
    from fusion_train import _split_k

    assert _split_k(128, 512, 16384) == 19          # 16 output tiles -> ~296 CTAs
    assert _split_k(4, 128, 1024) == 16             # capped by K / 64
    assert _split_k(4096, 4096, 64) == 1


def test_trainer_refuses_cpu_and_unbuilt_configurations():
    import b200_native as nat
    import model_module as mm
    import parameters_default as pd
    from fusion_train import FusionHeadTrainer

    params = pd.default_parameters()
    tr = FusionHeadTrainer(mm.FusionModel(params))
    assert len(tr.names) == 20 and tr.numel == sum(p.numel() for p in tr.params) and tr.flat_numel >= tr.numel
    with pytest.raises(nat.B200NativeError):
        tr.zero_grad()  # parameters live on the CPU: no CPU path
    params["fusion_model_parameters"]["fusion_specific_parameters"]["use_cross_attention"] = False
    params["fusion_model_parameters"]["use_se"] = False
    small = FusionHeadTrainer(mm.FusionModel(params), lambda_mask=0.2)   # gating + classifier + proj_in + mask head
    assert len(small.names) == 10 and not any("cross_attn" in n or "fusion_se" in n for n in small.names)
    with pytest.raises(ValueError):
        FusionHeadTrainer(mm.FusionModel(params), mask_loss_type="bce")


def test_shared_step_objective_selection_and_loud_cpu_failure():
    import b200_native as nat
    import model_module as mm
    import parameters_default as pd
    from train_fusion import LightningFusionModel

    params = pd.default_parameters()
    assert LightningFusionModel(None, None, mm.FusionModel(params), params)._lambda_mask() == 0.2
    params["fusion_model_parameters"].update(recon_enabled=True, lambda_recon=0.1, mimic_enabled=True, lambda_mimic=0.2)
    lm = LightningFusionModel(mm.ModelMaskHeadBackbone("dwi", params), mm.ModelMaskHeadBackbone("dce", params),
                              mm.FusionModel(params), params)
    obj = lm._objective()   # the reference's default objective is built in full ...
    assert obj == {"recon": True, "mimic": True, "lambda_recon": 0.1, "lambda_mimic": 0.2} and lm._needs_full_trainer()
    batch = (torch.zeros(2, 16, 64, 64), torch.zeros(2, 6, 64, 64), torch.zeros(2, 1, 32, 32), torch.zeros(2, dtype=torch.long))
    with pytest.raises(nat.B200NativeError):   # ... on the training kernels: no CPU path
        lm.training_step(batch)
    params["fusion_model_parameters"]["attn_reg_enabled"] = True
    with pytest.raises(NotImplementedError, match="attn_reg"):
        lm.training_step(batch)
    params["fusion_model_parameters"]["attn_reg_enabled"] = False
    # depth-wise parameter groups of LightningFusionOptimizerFactory (selector_helpers.py:456-518)
    params["fusion_model_parameters"]["optimizer_parameters"] = {
        "name": "adamW", "lr": 3e-4, "weight_decay": 4e-5, "discriminative_lr": True, "lr_decay_factor": 2.0,
        "discriminative_reg": True, "reg_decay_factor": 0.5, "reg_base": 1e-4}
    fn = lm._group_hparams()
    assert fn("fusion.classifier.2.weight") == (3e-4, 1e-4)
    assert fn("dwi.block3.skip.0.weight") == (1.5e-4, 5e-5) and fn("dce.mask_head.pre.weight") == (1.5e-4, 5e-5)
    assert fn("dwi.block2.se.fc.1.weight") == (7.5e-5, 2.5e-5) and fn("dce.block1.skip.1.bias") == (3.75e-5, 1.25e-5)
    # the token-shortcut head trainer serves the classification (+ mask) objective on eval-mode frozen encoders
    params["b200_frozen_encoder_mode"] = "eval"
    params["fusion_model_parameters"].update(recon_enabled=False, mimic_enabled=False)
    for m in (lm.dwi_model, lm.dce_model):
        for q in m.parameters():
            q.requires_grad = False
    assert not lm._needs_full_trainer()
    params["fusion_model_parameters"]["label_smoothing_enabled"] = False
    params["b200_classification_objective_only"] = True
    with pytest.raises(RuntimeError, match="label_smoothing"):
        lm.configure_optimizers()
    params["fusion_model_parameters"]["label_smoothing_enabled"] = True
    params["fusion_model_parameters"]["optimizer_parameters"] = {
        "name": "adamW", "lr": 3e-4, "betas": (0.9, 0.99), "eps": 1e-8, "weight_decay": 4e-5, "discriminative_lr": True,
        "lr_decay_factor": 1.2, "discriminative_reg": True, "reg_decay_factor": 0.8, "reg_base": 1e-4}
    tr = lm.configure_optimizers()   # last group of the discriminative schedule: base lr, reg_base weight decay
    assert (tr.lr, tr.weight_decay, tr.betas) == (3e-4, 1e-4, (0.9, 0.99))
    # the trainer is a torch.optim.Optimizer: the reference's scheduler choices (selector_helpers.py:692-728) drive it
    assert isinstance(tr, torch.optim.Optimizer) and len(tr.param_groups) == 1
    params["fusion_model_parameters"]["scheduler"] = {"name": "cosine", "T_max": 10, "eta_min": 1e-6}
    both = lm.configure_optimizers()
    tr, sch = both["optimizer"], both["lr_scheduler"]["scheduler"]
    twin = torch.optim.AdamW([torch.nn.Parameter(torch.zeros(1))], lr=3e-4)
    twin_sch = torch.optim.lr_scheduler.CosineAnnealingLR(twin, T_max=10, eta_min=1e-6)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")      # (scheduler.step() before optimizer.step(): no GPU here)
        for _ in range(4):
            sch.step()
            twin_sch.step()
    assert tr.lr == twin.param_groups[0]["lr"] and both["lr_scheduler"]["interval"] == "epoch"
    params["fusion_model_parameters"]["scheduler"] = {"name": "reduce_lr_on_plateau", "factor": 0.5, "patience": 0,
                                                      "min_lr": 1e-7, "threshold": 1e-4, "monitor": "val_loss"}
    both = lm.configure_optimizers()
    both["lr_scheduler"]["scheduler"].step(1.0)
    both["lr_scheduler"]["scheduler"].step(2.0)
    assert both["optimizer"].lr == 1.5e-4 and both["lr_scheduler"]["monitor"] == "val_loss"
    params["fusion_model_parameters"]["scheduler"] = {"name": "nope"}
    with pytest.raises(ValueError):
        lm.configure_optimizers()
    params["fusion_model_parameters"].pop("scheduler")
    w = lm.set_class_weights(torch.tensor([0, 0, 1, 2, 3, 3, 3, 3]))
    assert torch.allclose(w, torch.tensor([1.0, 2.0, 2.0, 0.5]), atol=1e-5)


_WORKER = r"""
import sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); import b200path
from fusion_train import average_gradients, flat_size, flat_views
rank = int(sys.argv[3])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=rank, world_size=2)
shapes = [torch.zeros(3, 2), torch.zeros(4)]
n = flat_size(shapes)
flat = torch.zeros(n + 1)                   # aligned gradient slots + the loss slot
g = flat_views(shapes, flat)
g[0].fill_(1.0 + rank); g[1].fill_(10.0 * (1 + rank)); flat[n] = 0.5 + rank
scale = average_gradients(flat)
assert scale == 0.5
avg = flat_views(shapes, flat * scale)
assert torch.allclose(avg[0], torch.full((3, 2), 1.5)) and torch.allclose(avg[1], torch.full((4,), 15.0))
assert abs(float(flat[n]) * scale - 1.0) < 1e-6     # the loss is averaged by the same collective
assert float((flat * scale).sum()) == 1.5 * 6 + 15.0 * 4 + 1.0   # the padding stays zero
dist.destroy_process_group()
print("ok")
"""


def test_gradient_allreduce_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(30100 + os.getpid() % 500)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


_BUCKET_WORKER = r"""
import sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); import b200path
from train_graph import FullFusionTrainer
rank = int(sys.argv[3])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=rank, world_size=2)
# the bucket plan of the all-unfrozen step on the host: 7 segments in forward order (3 per encoder + the fusion head),
# sizes in elements; the backward pass fires the markers last segment first
sizes = [900, 5000, 30000, 700, 5200, 31000, 4000]
ends, off = [], 0
for s in sizes:
    off += s; ends.append(off)
t = FullFusionTrainer.__new__(FullFusionTrainer)
t.dev, t.group, t._comm, t._pending = torch.device("cpu"), None, None, []
t.seg_ends, t.flat_numel, t.bucket_elems = ends, off, 8192
t.flat = {"g": torch.arange(off + 1, dtype=torch.float32) * (1 + rank)}      # + the loss slot, which no bucket covers
calls = []
orig = t._reduce_range
def spy(lo, hi):
    calls.append((lo, hi)); orig(lo, hi)
t._reduce_range = spy
t._reduced_from = t.flat_numel
for seg in reversed(range(len(sizes))):
    t._marker(seg)()
# every gradient element reduced exactly once, in backward order, buckets >= 8192 elements except the one that
# closes the buffer at segment 0
assert calls[0][1] == off and calls[-1][0] == 0 and all(a[0] == b[1] for a, b in zip(calls, calls[1:])), calls
assert all(hi - lo >= 8192 for lo, hi in calls[:-1]), calls
assert calls == [(41800, 76800), (5900, 41800), (0, 5900)], calls
want = torch.arange(off + 1, dtype=torch.float32) * 3.0                       # rank 0 + rank 1
want[off] = float(off) * (1 + rank)                                            # the loss slot stays local
assert torch.equal(t.flat["g"], want)
dist.destroy_process_group()
print("ok", calls)
"""


def test_full_trainer_bucket_plan_two_ranks_gloo(tmp_path):
    """FullFusionTrainer's gradient exchange (train_graph._marker / _reduce_range): tape markers fire in backward order,
    each closes a bucket once >= bucket_elems gradients are final, the first segment's marker flushes the rest; under
    gloo with two ranks every element of the flat gradient buffer is summed exactly once."""
    script = tmp_path / "bucket_worker.py"
    script.write_text(_BUCKET_WORKER)
    port = str(30700 + os.getpid() % 500)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_training_entry_points_reject_bad_arguments_without_a_gpu():
    import b200_native as nat

    lib = nat.lib()
    assert lib.b200_sgemm(None, 1, 0, None, 1, 0, None, 1, 4, 4, 4, None, None, 0, 1, None, 0, 0, 1, None) < 0
    assert lib.b200_sgemm(None, 1, 0, None, 1, 0, None, 1, 0, 4, 4, None, None, 0, 1, None, 0, 0, 1, None) == 0
    assert lib.b200_mha_fwd(None, 1, None, None, 1, 2, 4, 64, 16, 32, None, None, 1, None) < 0   # Tq > 32
    assert lib.b200_adamw(None, None, None, None, 8, 1e-3, 0.9, 0.999, 1e-8, 0.0, 0, 1.0, None) < 0  # step < 1
    assert lib.b200_head_loss(None, 4, None) < 0
    assert lib.b200_colsum(None, 1, 0, 4, None, None) == 0


def test_oracle_full_frozen_phase_objective_matches_the_reference():
    """classification + mask dice + fused reconstruction + mimic (train_fusion.py:238-292) with FusionModel in train
    mode (batch-statistic BatchNorm in ReconHead / Projector), computed by the reference's OWN loss functions
    (oracle/make_golden_train.py::main_full): the oracle restatement reproduces the total, every term and all 35
    gradients.  (The CUDA path builds the first two terms; this pins the oracle for the other two.)"""
    gold, hp, params, sd, batch = _setup("train_head_full.npz")
    dwi_in, dce_in, masks, _ = op.synthetic_raw(hp["n"], seed=hp["seed"] + 1, kind="S")
    dwi_in = dwi_in / dwi_in.amax(dim=(1, 2, 3), keepdim=True)
    loss, parts, grads = to.full_objective_and_grads(
        sd, params, *batch, masks, dwi_in, dce_in, hp["smoothing"], hp["gamma"], torch.tensor(hp["class_weights"]),
        hp["lambda_mask"], hp["lambda_recon"], hp["lambda_mimic"])
    total, cls, mask, recon, mimic = gold["parts"]
    assert abs(float(loss) - total) <= 1e-5 * abs(total)
    for got, want in ((parts["cls"], cls), (parts["mask"], mask), (parts["recon"], recon), (parts["mimic"], mimic)):
        assert abs(got - want) <= 1e-5 * abs(want), parts
    assert sorted(grads) == sorted(hp["with_grad"]) and len(grads) == 35
    for k in grads:
        gu.check(gold, f"grad/{k}", grads[k], 5e-4, what="gradient ")


def test_oracle_c1_single_modality_train_step_matches_the_reference():
    """BASELINE config C1 (DWI CNN forward + train step on the CPU): the reference's own modules and loss functions
    (oracle/make_golden_train.py::main_c1; train mode, dropout p = 0) against the oracle restatement - total, terms
    and all 89 parameter gradients.  (The CUDA path does not train encoders; this pins the oracle for it.)"""
    gold = gu.load("train_c1_dwi.npz")
    hp = json.loads(str(gold["hp"]))
    import parameters_default as pd

    params = pd.default_parameters()
    params["dwi_model_parameters"]["dropout"] = 0.0
    sd = op.seeded_state_dict(gu.load_shapes("cnn")["dwi"], seed=hp["weight_seed"])
    dwi_raw, _, masks, labels = op.synthetic_raw(hp["n"], seed=1234, kind="S")
    x = dwi_raw / dwi_raw.amax(dim=(1, 2, 3), keepdim=True)
    loss, parts, grads = to.single_model_objective_and_grads(
        sd, params, "dwi", x, masks, labels, hp["smoothing"], hp["gamma"], torch.tensor(hp["class_weights"]),
        hp["lambda_mask"], hp["lambda_recon"], hp["lambda_mimic"], hp["lambda_feat_norm"])
    total, cls, feat_norm, mask, recon_w, mimic_w = gold["parts"]
    assert abs(float(loss) - total) <= 2e-5 * abs(total)
    for got, want in ((parts["cls"], cls), (parts["feat_norm"], feat_norm), (parts["mask"], mask),
                      (parts["recon"] * hp["lambda_recon"], recon_w), (parts["mimic"] * hp["lambda_mimic"], mimic_w)):
        assert abs(got - want) <= 2e-5 * abs(want), (parts, gold["parts"])
    assert sorted(grads) == sorted(hp["with_grad"]) and len(grads) == 89
    worst = max(gu.check(gold, f"grad/{k}", grads[k], 1e-5, what="gradient ") for k in grads)
    assert worst <= 1e-5   # (same torch operators in the same order: bit-identical in the authoring container)
