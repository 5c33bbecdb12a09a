"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference, read-only) on seeded inputs and weights.  Run in the build container only:

    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so these fixtures — outputs
of the reference itself — are what pins the oracle (oracle/add_oracle.py) and, through it, the CUDA
path.  Weights are NOT stored: they are regenerated deterministically (CPU RNG, fixed seeds) by
`tests/util.py::make_weights`, which builds OUR drop-in module; loading its state_dict into the
reference module with strict=True is itself the key-compatibility check.  A checksum of the
weights is stored so RNG drift is detected instead of silently mis-compared.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, "/root/reference")

import util  # noqa: E402  (tests/util.py)
from modeling import operations as ref_ops  # noqa: E402
from modeling.ADD import ADD as RefADD, EDM as RefEDM, Cell as RefCell  # noqa: E402
from modeling.aspp_train import ASPP_train as RefASPP  # noqa: E402
from modeling.decoder import Decoder as RefDecoder  # noqa: E402
from utils.metrics import Evaluator as RefEvaluator  # noqa: E402

torch.cuda.synchronize = lambda *a, **k: None   # ADD.dynamic_inference calls it (ADD.py:380); CPU here
torch.set_num_threads(8)
OUT = Path(__file__).resolve().parent
BN = torch.nn.BatchNorm2d


def f32(t):
    return t.detach().cpu().numpy().astype(np.float32)


def gen_ops():
    g = {}
    for name, spec in util.OP_CASES.items():
        ours, x = util.make_op_case(name)
        kind, args = spec["kind"], spec["args"]
        if kind == "OPS":
            ref = ref_ops.OPS[args[0]](args[1], args[2] if len(args) > 2 else 1, BN, 1e-5, 0.1, True)
        else:
            ref = getattr(ref_ops, kind)(*args, BN)
        ref.load_state_dict(ours.state_dict(), strict=True)
        ref.eval()
        with torch.no_grad():
            y = ref(x)
        g[name + "/y"] = f32(y)
        g[name + "/wsum"] = np.float64(util.weight_checksum(ours.state_dict()))
    # ASPP_train / Decoder / EDM
    ours, x = util.make_aspp_case()
    ref = RefASPP(util.ASPP_CASE["C"], util.ASPP_CASE["out"], BN, depth=util.ASPP_CASE["depth"], mult=util.ASPP_CASE["mult"])
    ref.load_state_dict(ours.state_dict(), strict=True); ref.eval()
    with torch.no_grad():
        g["aspp/y"] = f32(ref(x))
    g["aspp/wsum"] = np.float64(util.weight_checksum(ours.state_dict()))
    ours, x, low, size = util.make_decoder_case()
    ref = RefDecoder(19, BN); ref.load_state_dict(ours.state_dict(), strict=True); ref.eval()
    with torch.no_grad():
        g["decoder/y"] = f32(ref(x.clone(), low.clone(), size))
    g["decoder/wsum"] = np.float64(util.weight_checksum(ours.state_dict()))
    ours, x = util.make_edm_case()
    ref = RefEDM(); ref.load_state_dict(ours.state_dict(), strict=True); ref.eval()
    with torch.no_grad():
        g["edm/y"] = f32(ref(x.clone()))
    g["edm/wsum"] = np.float64(util.weight_checksum(ours.state_dict()))
    # a Cell whose genotype uses every primitive (pools, skip_connect, none) — reference Cell, ADD.py:14-116
    ours, xpp, xp = util.make_cell_case()
    cc = util.CELL_CASE
    ref = RefCell(BN, 5, cc["prev_prev_C"], cc["prev_C"], util.MIXED_CELL.copy(), 1, cc["C_out"], 0, False, True)
    ref.load_state_dict(ours.state_dict(), strict=True); ref.eval()
    with torch.no_grad():
        _, concat, dense = ref(xpp, xp)
    g["cell_mixed/concat"] = f32(concat); g["cell_mixed/dense"] = f32(dense)
    g["cell_mixed/wsum"] = np.float64(util.weight_checksum(ours.state_dict()))
    # confidence scalars
    lg = util.make_logits_case()
    g["conf/entropy"] = np.float64(ref_ops.normalized_shannon_entropy(lg))
    g["conf/max_0.3"] = np.float64(ref_ops.confidence_max(lg, 0.3))
    g["conf/max_0.6"] = np.float64(ref_ops.confidence_max(lg, 0.6))
    # Evaluator
    for cname, (gt, pred) in util.make_evaluator_cases().items():
        ev = RefEvaluator(19)
        cm = ev._generate_matrix(gt, pred)
        g[f"evaluator/{cname}/cm"] = cm.numpy().astype(np.int64)
        ev.add_batch(gt, pred)
        g[f"evaluator/{cname}/miou"] = np.float64(ev.Mean_Intersection_over_Union())
    np.savez_compressed(OUT / "ops.npz", **g)
    print("ops.npz", len(g), "entries")


def gen_nets():
    g = {}
    for cname, spec in util.NET_CASES.items():
        ours = util.make_net(spec)
        na, ci, low = util.net_arch(spec)
        ref = RefADD(na, ci, util.cell_arch(), 19, SimpleNamespace(F=spec["F"], B=5, sync_bn=False), low)
        ref.load_state_dict(ours.state_dict(), strict=True)
        ref.eval()
        g[f"{cname}/wsum"] = np.float64(util.weight_checksum(ours.state_dict()))
        for (h, w) in spec["sizes"]:
            x, gt = util.make_input(1, h, w)
            with torch.no_grad():
                outs = ref(x)
            tag = f"{cname}/{h}x{w}"
            for e, o in enumerate(outs):
                g[f"{tag}/forward/{e}"] = f32(o)
                ev = RefEvaluator(19)
                g[f"{tag}/cm/{e}"] = ev._generate_matrix(gt, torch.argmax(o, 1)).numpy().astype(np.int64)
            if spec.get("dynamic"):
                with torch.no_grad():
                    lg, feat = ref.get_feature(x)
                g[f"{tag}/get_feature/logits"] = f32(lg)
                g[f"{tag}/get_feature/feature_sum"] = np.float64(feat.double().abs().sum().item())
                edm_ours = util.make_edm()
                edm = RefEDM(); edm.load_state_dict(edm_ours.state_dict(), strict=True); edm.eval()
                with torch.no_grad():
                    c0 = float(edm(feat.clone()))
                g[f"{tag}/edm_value"] = np.float64(c0)
                for label, thr in (("exit", c0 + 1.0), ("noexit", c0 - 1.0)):
                    with torch.no_grad():
                        y, ee, _, cv = ref.dynamic_inference(x, threshold=thr, confidence='edm', edm=edm)
                    assert ee == (1 if label == "exit" else 0)
                    g[f"{tag}/dynamic_edm/{label}/y"] = f32(y)
                    g[f"{tag}/dynamic_edm/{label}/conf"] = np.float64(float(cv))
        print(cname, "done")
    np.savez_compressed(OUT / "nets.npz", **g)
    print("nets.npz", len(g), "entries", os.path.getsize(OUT / "nets.npz") / 1e6, "MB")


def gen_siblings():
    """Baselin_Model (baseline_model.py:93-254) and AutoDeepLab (autodeeplab.py:94-204): the unmodified reference."""
    from modeling.baseline_model import Baselin_Model as RefBaseline
    from modeling.autodeeplab import AutoDeepLab as RefAutoDeepLab
    g = {}
    for cname, spec in util.SIBLING_CASES.items():
        ours = util.make_sibling(spec)
        na, ci, low = util.add_b200.NETWORKS[spec["network"]][spec["C"]]
        ns = SimpleNamespace(F=spec["F"], B=5, sync_bn=False)
        if spec["cls"] == "Baselin_Model":
            ref = RefBaseline(na, ci, util.cell_arch(), 19, ns, low)
        else:
            ref = RefAutoDeepLab(na, util.cell_arch(), 19, ns, low)
        ref.load_state_dict(ours.state_dict(), strict=True)
        ref.eval()
        x, _ = util.make_input(1, *spec["size"])
        with torch.no_grad():
            outs = ref(x)
        outs = outs if spec["cls"] == "Baselin_Model" else [outs[1]]
        for e, o in enumerate(outs):
            g[f"{cname}/forward/{e}"] = f32(o)
        g[f"{cname}/wsum"] = np.float64(util.weight_checksum(ours.state_dict()))
        print(cname, "done")
    np.savez_compressed(OUT / "siblings.npz", **g)
    print("siblings.npz", len(g), "entries")


def gen_syncbn():
    """SynchronizedBatchNorm2d in training mode, the unmodified reference class (batchnorm.py:180).  The DataParallel
    thread rendezvous (comm.py) is host control flow, not arithmetic: the fixture drives the reference's own pieces in
    the order `forward` / `_data_parallel_master` run them — per-shard sums (:59-61), sizes and sums added over the
    devices (:100-102; ReduceAddCoalesced is a plain sum), `_compute_mean_std` (:113-125, which also updates the running
    statistics) and the output expression (:68-75).  Also the one-device path (:50-53) through `forward` itself."""
    from modeling.sync_batchnorm.batchnorm import SynchronizedBatchNorm2d as RefSyncBN, _sum_ft, _unsqueeze_ft
    g = {}
    for cname, spec in util.SYNCBN_CASES.items():
        shards, state = util.make_syncbn_case(cname)
        C = spec["C"]
        for mode in ("sync", "local"):
            bn = RefSyncBN(C, eps=spec["eps"], momentum=spec["momentum"], affine=spec["affine"])
            bn.load_state_dict(state, strict=True)
            bn.train()
            with torch.no_grad():
                if mode == "local":
                    outs = [bn(x) for x in shards]                      # not parallel -> F.batch_norm (:50-53)
                else:
                    flat = [x.view(x.size(0), C, -1) for x in shards]
                    size = sum(f.size(0) * f.size(2) for f in flat)
                    sum_ = sum(_sum_ft(f) for f in flat)
                    ssum = sum(_sum_ft(f ** 2) for f in flat)
                    mean, inv_std = bn._compute_mean_std(sum_, ssum, size)
                    outs = []
                    for x, f in zip(shards, flat):
                        if bn.affine:
                            o = (f - _unsqueeze_ft(mean)) * _unsqueeze_ft(inv_std * bn.weight) + _unsqueeze_ft(bn.bias)
                        else:
                            o = (f - _unsqueeze_ft(mean)) * _unsqueeze_ft(inv_std)
                        outs.append(o.view(x.size()))
                    g[f"{cname}/{mode}/mean"] = f32(mean)
                    g[f"{cname}/{mode}/inv_std"] = f32(inv_std)
            for i, o in enumerate(outs):
                g[f"{cname}/{mode}/y{i}"] = f32(o)
            g[f"{cname}/{mode}/running_mean"] = f32(bn.running_mean)
            g[f"{cname}/{mode}/running_var"] = f32(bn.running_var)
    np.savez_compressed(OUT / "syncbn.npz", **g)
    print("syncbn.npz", len(g), "arrays")


def gen_train_ops():
    """The conv operators in TRAINING mode (module.train(): batch statistics, running statistics updated), the
    unmodified reference modules on one device."""
    g = {}
    for name in util.TRAIN_OP_CASES:
        spec = util.OP_CASES[name]
        ours, x = util.make_op_case(name)
        kind, args = spec["kind"], spec["args"]
        if kind == "OPS":
            ref = ref_ops.OPS[args[0]](args[1], args[2] if len(args) > 2 else 1, BN, 1e-5, 0.1, True)
        else:
            ref = getattr(ref_ops, kind)(*args, BN)
        ref.load_state_dict(ours.state_dict(), strict=True)
        ref.train()
        with torch.no_grad():
            y = ref(x)
        g[name + "/y"] = f32(y)
        for k, v in ref.state_dict().items():
            if "running_" in k:
                g[f"{name}/sd/{k}"] = f32(v)
    np.savez_compressed(OUT / "train_ops.npz", **g)
    print("train_ops.npz", len(g), "arrays")


def gen_mixed_op():
    """MixedOp (cell_level_search.py:10-29), the unmodified reference: weighted sum in eval and in training mode
    (batch statistics, running statistics updated) and the argmax path (`training=False`)."""
    from modeling.cell_level_search import MixedOp as RefMixedOp
    g = {}
    for name, spec in util.MIXED_OP_CASES.items():
        ours, x, w = util.make_mixed_op_case(name)
        for mode in ("eval", "train"):
            ref = RefMixedOp(spec["C"], 1, BN)
            ref.load_state_dict(ours.state_dict(), strict=True)
            ref.train(mode == "train")
            with torch.no_grad():
                g[f"{name}/{mode}/y"] = f32(ref(x, w))
                if mode == "eval":
                    g[f"{name}/eval/y_argmax"] = f32(ref(x, w, training=False))
            if mode == "train":
                for k, v in ref.state_dict().items():
                    if "running_" in k:
                        g[f"{name}/train/sd/{k}"] = f32(v)
        g[name + "/w"] = f32(w)
    np.savez_compressed(OUT / "mixed_op.npz", **g)
    print("mixed_op.npz", len(g), "arrays")


def gen_gates():
    """`dynamic_inference(confidence='entropy' | 'max')` (ADD.py:440-488), the unmodified reference.  The reference
    returns the feature map `x` there instead of the logits (:488); what the exit head produced is recorded through a
    forward hook on `ref.decoder` (observation only).  Stored per gate and per decision: the earlier_exit flag, the
    confidence value, the logits of the LAST decoder call (= the exit that was taken) and |x|-sum of the returned map."""
    g = {}
    spec = util.NET_CASES["searched-dense-C2"]
    ours = util.make_net(spec)
    na, ci, low = util.net_arch(spec)
    ref = RefADD(na, ci, util.cell_arch(), 19, SimpleNamespace(F=spec["F"], B=5, sync_bn=False), low)
    ref.load_state_dict(ours.state_dict(), strict=True)
    ref.eval()
    seen = []
    ref.decoder.register_forward_hook(lambda m, i, o: seen.append(o.detach().clone()))
    for (h, w) in spec["sizes"]:
        x, _ = util.make_input(1, h, w)
        tag = f"searched-dense-C2/{h}x{w}"
        # entropy of the first exit's logits decides; thresholds on either side of it
        with torch.no_grad():
            seen.clear()
            _, _, _, e0 = ref.dynamic_inference(x, threshold=-1.0, confidence='entropy')      # never exits: e0 = exit-1 entropy
        g[f"{tag}/entropy/value"] = np.float64(e0)
        cases = [("entropy", "exit", float(e0) + 0.05), ("entropy", "noexit", float(e0) - 0.05),
                 ("max", "exit", 0.05), ("max", "noexit", 1.0)]
        for conf, label, thr in cases:
            seen.clear()
            with torch.no_grad():
                xret, ee, _, cv = ref.dynamic_inference(x, threshold=thr, confidence=conf)
            assert ee == (1 if label == "exit" else 0), (conf, label, thr, ee, cv)
            k = f"{tag}/{conf}/{label}"
            g[k + "/threshold"] = np.float64(thr)
            g[k + "/conf"] = np.float64(float(cv))
            g[k + "/y"] = f32(seen[-1])
            g[k + "/n_heads"] = np.int64(len(seen))
            g[k + "/x_abs_sum"] = np.float64(xret.double().abs().sum().item())
    g["searched-dense-C2/wsum"] = np.float64(util.weight_checksum(ours.state_dict()))
    np.savez_compressed(OUT / "gates.npz", **g)
    print("gates.npz", len(g), "entries", os.path.getsize(OUT / "gates.npz") / 1e6, "MB")


def gen_three_gates():
    """`dynamic_inference(confidence='edm')` through a network with THREE gated exits (ADD.py:394-438: the loop keeps going
    while edm(y) > threshold; conv_aspp_iter counts the skipped exits; EDM's in-place ReLU is seen by every later cell),
    the unmodified reference, image by image (its gate is batch-1).  Thresholds are derived from the reference's own
    gate values: between the 2nd/3rd and the 4th/5th smallest first-gate values, never-exit, and the median of the
    last-gate values (images leave at gate 3 or run to the end).  Stored for every (threshold, image): earlier_exit,
    the confidence value returned, sum|y| and the argmax histogram; the logits themselves for images 0 and 3."""
    g = {}
    ours, edm, x, _ = util.make_three_gate_case()
    c = util.THREE_GATES
    ref = RefADD(c["network_arch"], c["C_index"], util.cell_arch(), 19, SimpleNamespace(F=c["F"], B=c["B"], sync_bn=False),
                 c["low_level_layer"])
    ref.load_state_dict(ours.state_dict(), strict=True)
    ref.eval()
    redm = RefEDM()
    redm.load_state_dict(edm.state_dict(), strict=True)
    redm.eval()

    def run(i, thr):
        with torch.no_grad():
            y, ee, _, cv = ref.dynamic_inference(x[i:i + 1].clone(), threshold=thr, confidence='edm', edm=redm)
        return y, int(ee), float(cv)
    n = c["n"]
    firsts = [run(i, 1e30)[2] for i in range(n)]            # exits at the first gate: its value
    lasts = [run(i, -1e30)[2] for i in range(n)]            # never exits: the last gate's value
    srt = sorted(firsts)
    thresholds = [0.5 * (srt[1] + srt[2]), -1e30, float(np.median(lasts)), 0.5 * (srt[3] + srt[4])]
    g["thresholds"] = np.array(thresholds, dtype=np.float64)
    g["first_gate_values"] = np.array(firsts, dtype=np.float64)
    g["last_gate_values"] = np.array(lasts, dtype=np.float64)
    for t, thr in enumerate(thresholds):
        for i in range(n):
            y, ee, cv = run(i, thr)
            k = f"t{t}/img{i}"
            g[k + "/exit"] = np.int64(ee)
            g[k + "/conf"] = np.float64(cv)
            g[k + "/y_abs_sum"] = np.float64(y.double().abs().sum().item())
            g[k + "/argmax_hist"] = np.bincount(y.argmax(1).flatten().numpy(), minlength=19).astype(np.int64)
            if i in (0, 3):
                g[k + "/y"] = f32(y)
    g["wsum"] = np.float64(util.weight_checksum(ours.state_dict()))
    np.savez_compressed(OUT / "three_gates.npz", **g)
    print("three_gates.npz", len(g), "entries", os.path.getsize(OUT / "three_gates.npz") / 1e6, "MB",
          "exit flags per threshold:", [[int(g[f"t{t}/img{i}/exit"]) for i in range(n)] for t in range(len(thresholds))])


def gen_io_edges():
    """Loader / dump edges, the unmodified reference: CityscapesSegmentation.encode_segmap (cityscapes.py:85-91; the
    class is constructed with its file glob stubbed — the dataset is not here), full_image_eval_preprocess
    (custom_transforms.py:322-347) on PIL inputs, decode_segmap (dataloaders/utils.py:14-51; matplotlib, which that
    module imports for its optional plot, is absent here and stubbed)."""
    import types
    from PIL import Image
    sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
    sys.modules.setdefault("matplotlib.pyplot", types.ModuleType("matplotlib.pyplot"))
    from dataloaders.datasets.cityscapes import CityscapesSegmentation
    from dataloaders import custom_transforms as tr
    from dataloaders.utils import decode_segmap as ref_decode
    CityscapesSegmentation.recursive_glob = lambda self, rootdir='.', suffix='': ['stub.png']
    ds = CityscapesSegmentation(None, root='/tmp', split='val')
    g = {}
    g["encode/all_ids"] = ds.encode_segmap(np.arange(256, dtype=np.uint8).reshape(16, 16).copy())
    for name, spec in util.IO_CASES.items():
        img, ids = util.make_io_case(name)
        enc = ds.encode_segmap(ids.copy())
        g[f"{name}/encoded"] = enc
        out = tr.full_image_eval_preprocess(spec["crop"], ds.mean, ds.std)({'image': Image.fromarray(img), 'label': Image.fromarray(enc)})
        g[f"{name}/image"] = f32(out['image'])
        g[f"{name}/label"] = out['label'].numpy().astype(np.int64)
        g[f"{name}/decoded"] = ref_decode(enc.astype(np.int64), 'cityscapes')
    np.savez_compressed(OUT / "io_edges.npz", **g)
    print("io_edges.npz", len(g), "arrays")


def gen_search_cell():
    """The supernet cell (cell_level_search.py:32-155) and the search-time ASPP head (operations.py:122-158), the
    unmodified reference: eval-mode forward, training-mode forward + backward (loss = sum of outputs x cotangents) with
    the alphas softmaxed as in model_net_search.py:294-310 — gradients of the inputs, the raw alphas and every weight,
    and the running statistics after the training forward."""
    from modeling.cell_level_search import Cell as RefSearchCell
    g = {}
    c = util.SEARCH_CELL
    ours, s0, s1_same, s1_up, alphas, cots = util.make_search_cell_case()
    ref = RefSearchCell(c["B"], c["prev_prev_C"], c["prev_C_down"], c["prev_C_same"], c["prev_C_up"], c["C_out"])
    ref.load_state_dict(ours.state_dict(), strict=True)
    ref.eval()
    with torch.no_grad():
        outs = ref(s0.clone(), None, s1_same.clone(), s1_up.clone(), torch.softmax(alphas, dim=-1))
    for i, o in enumerate(outs):
        g[f"cell/eval/out{i}"] = f32(o)
    ref.train()
    ins = [t.clone().requires_grad_(True) for t in (s0, s1_same, s1_up)]
    al = alphas.clone().requires_grad_(True)
    outs = ref(ins[0], None, ins[1], ins[2], torch.softmax(al, dim=-1))
    sum((o * ct).sum() for o, ct in zip(outs, cots)).backward()
    for i, o in enumerate(outs):
        g[f"cell/train/out{i}"] = f32(o)
    for nme, t in zip(("s0", "s1_same", "s1_up"), ins):
        g[f"cell/train/d_{nme}"] = f32(t.grad)
    g["cell/train/d_alphas"] = f32(al.grad)
    for k, p_ in ref.named_parameters():
        g[f"cell/train/grad/{k}"] = f32(p_.grad)
    for k, v in ref.state_dict().items():
        if "running_" in k:
            g[f"cell/train/sd/{k}"] = f32(v)
    # search-time ASPP
    a = util.SEARCH_ASPP
    ours, x, cot = util.make_search_aspp_case()
    ref = ref_ops.ASPP(a["C"], a["out"], a["pad"], a["dil"])
    ref.load_state_dict(ours.state_dict(), strict=True)
    ref.eval()
    with torch.no_grad():
        g["aspp/eval/y"] = f32(ref(x.clone()))
    ref.train()
    xr = x.clone().requires_grad_(True)
    y = ref(xr)
    (y * cot).sum().backward()
    g["aspp/train/y"] = f32(y)
    g["aspp/train/dx"] = f32(xr.grad)
    for k, p_ in ref.named_parameters():
        g[f"aspp/train/grad/{k}"] = f32(p_.grad)
    np.savez_compressed(OUT / "search_cell.npz", **g)
    print("search_cell.npz", len(g), "arrays", os.path.getsize(OUT / "search_cell.npz") / 1e6, "MB")


def gen_train_ops_grad():
    """Backward of the conv operators in TRAINING mode: the unmodified reference modules, loss = sum(y * cotangent),
    `.backward()` -> input gradient and every parameter gradient."""
    g = {}
    for name in util.TRAIN_OP_CASES:
        spec = util.OP_CASES[name]
        ours, x, cot = util.make_op_grad_case(name)
        kind, args = spec["kind"], spec["args"]
        if kind == "OPS":
            ref = ref_ops.OPS[args[0]](args[1], args[2] if len(args) > 2 else 1, BN, 1e-5, 0.1, True)
        else:
            ref = getattr(ref_ops, kind)(*args, BN)
        ref.load_state_dict(ours.state_dict(), strict=True)
        ref.train()
        xr = x.clone().requires_grad_(True)
        y = ref(xr)
        (y * cot).sum().backward()
        g[f"{name}/y"] = f32(y)
        g[f"{name}/dx"] = f32(xr.grad)
        for k, p_ in ref.named_parameters():
            g[f"{name}/grad/{k}"] = f32(p_.grad)
    np.savez_compressed(OUT / "train_ops_grad.npz", **g)
    print("train_ops_grad.npz", len(g), "arrays", os.path.getsize(OUT / "train_ops_grad.npz") / 1e6, "MB")


def gen_train_step():
    """Two iterations of train.py:216-247 with the unmodified reference ADD in train mode: forward (batch-statistics
    BatchNorm), loss = mean over the exits of nn.CrossEntropyLoss(ignore_index=255) (utils/loss.py), backward,
    torch.optim.SGD(momentum, weight_decay, nesterov) (train.py:126-127).  Stored: the losses, per parameter the sum and
    abs-sum of its first-step gradient (full tensors for a few), and per parameter / running statistic the sum and abs-sum
    after each step."""
    spec = util.TRAIN_STEP
    ours, x, gt = util.make_train_case()
    na, ci, low = util.net_arch(spec)
    ref = RefADD(na, ci, util.cell_arch(), 19, SimpleNamespace(F=spec["F"], B=5, sync_bn=False), low)
    ref.load_state_dict(ours.state_dict(), strict=True)
    ref.train()
    opt = torch.optim.SGD(ref.parameters(), lr=spec["lr"], momentum=spec["momentum"], weight_decay=spec["weight_decay"],
                          nesterov=spec["nesterov"])
    crit = torch.nn.CrossEntropyLoss(ignore_index=255)
    g = {}
    for step in range(spec["steps"]):
        opt.zero_grad()
        outs = ref(x)
        losses = [crit(o, gt) for o in outs]
        loss = sum(losses) / len(losses)
        loss.backward()
        g[f"step{step}/loss"] = np.float64(loss.item())
        g[f"step{step}/exit_losses"] = np.array([l.item() for l in losses], dtype=np.float64)
        if step == 0:
            names, sums, asums = [], [], []
            for k, p_ in ref.named_parameters():
                names.append(k); sums.append(p_.grad.double().sum().item()); asums.append(p_.grad.double().abs().sum().item())
                if k in util.TRAIN_FULL_GRADS:
                    g[f"grad/{k}"] = f32(p_.grad)
            g["grad_names"] = np.array(names)
            g["grad_sum"] = np.array(sums, dtype=np.float64)
            g["grad_abs_sum"] = np.array(asums, dtype=np.float64)
            # conditioning of this gradient: the reference's OWN first-step gradients when the input is perturbed by 1e-6
            # relative (a random-init 12-cell network with BatchNorm over small batches is chaotic: they move by percents)
            import copy
            ref_p = copy.deepcopy(ref)
            ref_p.zero_grad()
            xp = x * (1 + 1e-6 * torch.randn(x.shape, generator=torch.Generator().manual_seed(62)))
            outs_p = ref_p(xp)
            (sum(crit(o, gt) for o in outs_p) / len(outs_p)).backward()
            g["grad_abs_sum_pert"] = np.array([p_.grad.double().abs().sum().item() for _, p_ in ref_p.named_parameters()], dtype=np.float64)
            g["grad_rel_change_pert"] = np.array([float((p_.grad - q_.grad).abs().max() / q_.grad.abs().max().clamp_min(1e-30))
                                                  for (_, p_), (_, q_) in zip(ref_p.named_parameters(), ref.named_parameters())], dtype=np.float64)
        opt.step()
        sd = ref.state_dict()
        keys = [k for k, v in sd.items() if v.dtype.is_floating_point]
        g[f"step{step}/state_names"] = np.array(keys)
        g[f"step{step}/state_sum"] = np.array([sd[k].double().sum().item() for k in keys], dtype=np.float64)
        g[f"step{step}/state_abs_sum"] = np.array([sd[k].double().abs().sum().item() for k in keys], dtype=np.float64)
    np.savez_compressed(OUT / "train_step.npz", **g)
    print("train_step.npz", len(g), "arrays", os.path.getsize(OUT / "train_step.npz") / 1e6, "MB", "losses", g["step0/loss"], g["step1/loss"])


if __name__ == "__main__":
    if "--only-search" in sys.argv:
        gen_search_cell()
        sys.exit(0)
    if "--only-train" in sys.argv:
        gen_train_ops_grad()
        gen_train_step()
        sys.exit(0)
    if "--only-io" in sys.argv:
        gen_io_edges()
        sys.exit(0)
    if "--only-gates" in sys.argv:
        gen_gates()
        gen_three_gates()
        sys.exit(0)
    if "--only-three-gates" in sys.argv:
        gen_three_gates()
        sys.exit(0)
    if "--only-syncbn" in sys.argv:
        gen_syncbn()
        gen_train_ops()
        gen_mixed_op()
        sys.exit(0)
    gen_ops()
    gen_nets()
    gen_siblings()
    gen_syncbn()
    gen_train_ops()
    gen_mixed_op()
    gen_gates()
    gen_three_gates()
    gen_io_edges()
    gen_train_step()
    gen_train_ops_grad()
    gen_search_cell()
