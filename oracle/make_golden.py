"""Generate tests/golden/*.npz by running the UNMODIFIED reference functions.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden

The reference has no tests or golden vectors of its own, so
parity is pinned by recording what its functions return on seeded inputs.
Inputs are stored next to the outputs, so the fixtures do not depend on RNG
reproducibility across machines.  Environment at generation time is recorded
in tests/golden/MANIFEST.json.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def softmax_slab(gen, P, C, spatial, scale):
    return torch.softmax(scale * torch.randn(P, C, *spatial, generator=gen), dim=1)


def uncertainty_cases(ref):
    gen = torch.Generator().manual_seed(20260101)
    cases = {}
    specs = [
        # name, P, C, spatial, logit scale
        ("cfg1_toy2d", 10, 2, (16, 32), 2.0),
        ("cfg2_lidc3d", 5, 2, (8, 8, 8), 2.0),
        ("cfg3_gta", 10, 19, (8, 16), 2.0),
        ("cfg3_gta_peaked", 10, 19, (8, 16), 8.0),
        ("cfg4_diffusion", 32, 2, (16, 16), 2.0),
        ("cfg4_diffusion_peaked", 32, 2, (16, 16), 12.0),
        ("cfg5_sweep", 16, 19, (8, 16), 3.0),
        ("p17_c3", 17, 3, (8, 16), 2.0),
        ("p18_c3", 18, 3, (8, 16), 2.0),
        ("p33_c4", 33, 4, (8, 8), 2.0),
        ("p2_c7_odd", 2, 7, (5, 13), 2.0),
    ]
    for name, P, C, spatial, scale in specs:
        cases[name] = softmax_slab(gen, P, C, spatial, scale)

    # edge vectors
    P, C, S = 6, 4, (4, 16)
    onehot = torch.zeros(P, C, *S)
    idx = torch.randint(0, C, (P, *S), generator=gen)
    onehot.scatter_(1, idx.unsqueeze(1), 1.0)
    cases["edge_onehot_ties"] = onehot  # --discretize: one-hot members, tied means
    same = softmax_slab(gen, 1, C, S, 2.0).repeat(P, 1, 1, 1)
    cases["edge_identical_members"] = same  # EU ~ 0, slightly negative allowed
    zeros = torch.zeros(P, C, *S)
    cases["edge_zeros"] = zeros
    mixed = softmax_slab(gen, P, C, S, 4.0)
    mixed[0, 0, 0, :4] = float("nan")
    mixed[1, 1, 1, :4] = -0.25
    mixed[2, 2, 2, :4] = 0.0
    mixed[3, 3, 3, :4] = 1e-42  # subnormal
    mixed[4, 0, 3, 8:12] = 1.0
    cases["edge_nan_neg_zero_subnormal"] = mixed
    near1 = torch.full((P, 2, *S), 0.0)
    eps = torch.logspace(-8, -1, S[0] * S[1]).reshape(S)
    near1[:, 0] = 1.0 - eps
    near1[:, 1] = eps
    near1[1::2, 0] = (1.0 - 0.5 * eps)
    near1[1::2, 1] = 0.5 * eps
    cases["edge_near_one"] = near1  # confident pixels: log accuracy near p = 1
    out = {}
    for name, x in cases.items():
        with torch.no_grad():
            u = ref.calculate_uncertainty(x)
        m = torch.mean(x, dim=0)
        out[f"{name}/x"] = x.numpy()
        out[f"{name}/TU"] = u["TU"].numpy()
        out[f"{name}/AU"] = u["AU"].numpy()
        out[f"{name}/EU"] = u["EU"].numpy()
        out[f"{name}/mean"] = m.numpy()
        out[f"{name}/label"] = m.argmax(dim=0).numpy().astype(np.uint8)
    # P == 1 -> one-minus-MSR (test_2D.py:1006-1007)
    x1 = softmax_slab(gen, 1, 5, (8, 16), 3.0)
    out["msr_p1/x"] = x1.numpy()
    out["msr_p1/pred_entropy"] = ref.calculate_one_minus_msr(x1.squeeze(0))["pred_entropy"].numpy()
    out["msr_p1/label"] = x1[0].argmax(dim=0).numpy().astype(np.uint8)
    return out


def aggregation_cases(ref):
    rng = np.random.default_rng(7)
    out = {}
    img2d = (rng.random((40, 56)) ** 3).astype(np.float32)
    img3d = (rng.random((14, 16, 18)) ** 3).astype(np.float32)
    plateau = np.zeros((32, 32), np.float32)
    plateau[5:20, 7:25] = 0.5  # many boxes tie for the max -> first-isclose rule
    for name, img in (("img2d", img2d), ("img3d", img3d), ("plateau", plateau)):
        out[f"{name}/image"] = img
        out[f"{name}/image_level_mean"] = np.float64(ref.image_level_aggregation(img)["max_score"])
        out[f"{name}/image_level_sum"] = np.float64(ref.image_level_aggregation(img, mean=False)["max_score"])
        for ps in (10, 4):
            r = ref.patch_level_aggregation(img, ps)
            out[f"{name}/patch{ps}_score"] = np.float64(r["max_score"])
            out[f"{name}/patch{ps}_bbox"] = np.asarray(r["bounding_box"], dtype=np.int64)
            r = ref.patch_level_aggregation(img, ps, mean=True)
            out[f"{name}/patch{ps}_mean_score"] = np.float64(r["max_score"])
        for tname, t in (("mid", float(np.quantile(img, 0.9))), ("above_max", float(img.max()) + 1.0),
                         ("zero", 0.0)):
            r = ref.threshold_aggregation(img, threshold=t)
            out[f"{name}/thr_{tname}_t"] = np.float64(t)
            out[f"{name}/thr_{tname}_score"] = np.float64(r["max_score"])
            r = ref.threshold_aggregation(img, threshold=t, mean=False)
            out[f"{name}/thr_{tname}_sum"] = np.float64(r["max_score"])
    # area / border on label maps
    lab2d = (rng.random((24, 40)) > 0.7).astype(np.uint8) * rng.integers(1, 19, (24, 40)).astype(np.uint8)
    lab3d = (rng.random((6, 10, 12)) > 0.5).astype(np.uint8)
    empty = np.zeros((8, 8), np.uint8)
    for name, lab in (("lab2d", lab2d), ("lab3d", lab3d), ("empty", empty)):
        out[f"{name}/label"] = lab
        out[f"{name}/area"] = np.float64(ref.compute_area(lab))
        out[f"{name}/border"] = np.float64(ref.compute_border(lab))
        out[f"{name}/norm_area"] = np.float64(ref.normalize_uncertainty_sum(img2d, ref.compute_area(lab)))
    return out


def calibration_cases(ref):
    rng = np.random.default_rng(11)
    out = {}
    H, W, R = 24, 32, 4
    specs = {
        "lidc_like": dict(n_cls=2, ignore=None, a=3.5, b=-1.25),
        "gta_like_ignore": dict(n_cls=19, ignore=255, a=6.0, b=-2.0),
        "steep": dict(n_cls=2, ignore=None, a=900.0, b=-300.0),  # many edges collapse onto few floats
        "negative_a": dict(n_cls=2, ignore=None, a=-2.0, b=0.5),
        "all_correct": dict(n_cls=1, ignore=None, a=3.5, b=-1.25),  # single-class quirk Q8
    }
    for name, s in specs.items():
        pred = rng.integers(0, s["n_cls"], (H, W)).astype(np.uint8)
        refs = np.stack([np.where(rng.random((H, W)) < 0.8, pred, rng.integers(0, max(s["n_cls"], 2), (H, W)))
                         for _ in range(R)]).astype(np.uint8)
        if name == "all_correct":
            refs = np.stack([pred] * R)
        if s["ignore"] is not None:
            refs[rng.random(refs.shape) < 0.05] = s["ignore"]
        unc = (rng.random((H, W)) ** 2 * 0.69).astype(np.float32)
        unc[0, :4] = 0.0
        with tempfile.TemporaryDirectory() as td:
            pf = os.path.join(td, "platt_scale_params.json")
            with open(pf, "w") as f:
                json.dump({"TU": {"a": s["a"], "b": s["b"]}}, f)
            # the body of calibration_error (ace.py:484-515), calling the reference pieces
            n_gt = refs.shape[0]
            pred_rep = np.repeat(pred[np.newaxis, :], n_gt, 0)
            unc_rep = np.repeat(unc[np.newaxis, :], n_gt, 0)
            correct = (refs == pred_rep).astype(int)
            if s["ignore"] is not None:
                keep = refs != s["ignore"]
                conf = ref.platt_scale_confid(-unc_rep[keep], platt_scale_file=pf, uncertainty="TU")
                cv = correct[keep]
            else:
                conf = ref.platt_scale_confid(-unc_rep.flatten(), platt_scale_file=pf, uncertainty="TU")
                cv = correct.flatten()
        acc = ref.GlobalCalibAccumulator()
        acc.accumulate(cv, conf)
        out[f"{name}/refs"] = refs
        out[f"{name}/pred"] = pred
        out[f"{name}/unc"] = unc
        out[f"{name}/a"] = np.float64(s["a"])
        out[f"{name}/b"] = np.float64(s["b"])
        out[f"{name}/ignore"] = np.int64(-999 if s["ignore"] is None else s["ignore"])
        out[f"{name}/conf"] = conf
        out[f"{name}/correct"] = cv.astype(np.uint8)
        out[f"{name}/ace"] = np.float64(ref.calc_ace(cv, conf))
        out[f"{name}/ece"] = np.float64(ref.calc_ece(cv, conf))
        out[f"{name}/g_bin_sums"] = acc.bin_sums
        out[f"{name}/g_bin_true"] = acc.bin_true
        out[f"{name}/g_bin_total"] = acc.bin_total
        out[f"{name}/gace"] = np.float64(acc.compute_ace())
        out[f"{name}/gece"] = np.float64(acc.compute_ece())
    return out


def ncc_aurc_cases(ref):
    rng = np.random.default_rng(13)
    out = {}
    H, W, R = 32, 32, 4
    refs = (rng.random((R, H, W)) < np.linspace(0.1, 0.9, W)[None, None, :]).astype(np.uint8)
    pred = (rng.random((H, W)) * 0.69).astype(np.float32)
    gt_map = np.var(refs, axis=0)  # experiment_dataloader.py:283
    out["ncc/refs"] = refs
    out["ncc/pred"] = pred
    out["ncc/gt_map"] = gt_map
    out["ncc/value"] = np.float64(ref.compute_ncc(gt_map, pred))
    out["ncc/self"] = np.float64(ref.compute_ncc(pred, pred))
    out["ncc/const_gt"] = np.float64(ref.compute_ncc(np.zeros((H, W)), pred))
    out["ncc/const_pred"] = np.float64(ref.compute_ncc(gt_map, np.full((H, W), 0.25, np.float32)))
    n = 300
    risks = rng.random(n)
    confids = -rng.random(n)
    out["aurc/risks"] = risks
    out["aurc/confids"] = confids
    out["aurc/aurc"] = np.float64(ref.aurc(risks, confids))
    out["aurc/eaurc"] = np.float64(ref.eaurc(risks, confids))
    cov, sel, w = ref.rc_curve_stats(risks, confids)
    out["aurc/coverages"] = np.asarray(cov, np.float64)
    out["aurc/selective_risks"] = np.asarray(sel, np.float64)
    out["aurc/weights"] = np.asarray(w, np.float64)
    return out


def platt_fit_cases(ref):
    """The reference's platt_scale_params (ace.py:14-285) run unmodified on a small in-memory validation set; its call to
    sklearn's fit is observed (not replaced) to record the compressed per-bin samples it builds."""
    import pathlib
    import types
    rng = np.random.default_rng(17)
    out = {}
    specs = {"lidc_like": dict(n_cls=2, ignore=None, R=4, n_img=3), "gta_like_ignore": dict(n_cls=19, ignore=255, R=2, n_img=2)}
    H, W = 24, 40
    for name, s in specs.items():
        images = []
        for i in range(s["n_img"]):
            pred = rng.integers(0, s["n_cls"], (H, W)).astype(np.uint8)
            maps = {}
            for unc in ("TU", "AU", "EU"):
                m = (10.0 ** rng.uniform(-9, -0.2, (H, W))).astype(np.float32)
                m[0, :5] = 0.0            # below the first edge: clamped into bin 0
                m[1, :3] = 150.0          # above the last edge: clamped into bin 255
                m[2, 0] = np.float32(1e-12)
                maps[unc] = m
            wrong = rng.random((s["R"], H, W)) < np.clip(maps["TU"] * 3, 0.02, 0.9)[None]
            refs = np.where(wrong, rng.integers(0, max(s["n_cls"], 2), (s["R"], H, W)), pred[None]).astype(np.uint8)
            if s["ignore"] is not None:
                refs[rng.random(refs.shape) < 0.05] = s["ignore"]
            images.append((refs, pred, maps))
        captured = {}
        real_calib = ref.ace_module.calib

        def spy(F, y, sample_weight=None, _store=captured, _real=real_calib):
            _store.setdefault("calls", []).append((np.array(F), np.array(y), np.array(sample_weight)))
            return _real(F, y, sample_weight=sample_weight)

        with tempfile.TemporaryDirectory() as td:
            loader = types.SimpleNamespace(
                exp_version=types.SimpleNamespace(unc_types=["TU", "AU", "EU"], exp_path=pathlib.Path(td)),
                image_ids=list(range(s["n_img"])),
                get_reference_segs=lambda i: images[i][0], get_mean_pred_seg=lambda i: images[i][1],
                get_unc_map=lambda i, unc: images[i][2][unc])
            ref.ace_module.calib = spy
            try:
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    ref.platt_scale_params(loader, ignore_value=s["ignore"])
            finally:
                ref.ace_module.calib = real_calib
            with open(os.path.join(td, "platt_scale_params.json")) as f:
                params = json.load(f)
        out[f"{name}/n_img"] = np.int64(s["n_img"])
        out[f"{name}/ignore"] = np.int64(-999 if s["ignore"] is None else s["ignore"])
        for i, (refs, pred, maps) in enumerate(images):
            out[f"{name}/refs{i}"] = refs
            out[f"{name}/pred{i}"] = pred
            for unc in ("TU", "AU", "EU"):
                out[f"{name}/{unc}{i}"] = maps[unc]
        for k, unc in enumerate(("TU", "AU", "EU")):
            F, y, w = captured["calls"][k]
            out[f"{name}/{unc}_F"] = F
            out[f"{name}/{unc}_y"] = y
            out[f"{name}/{unc}_w"] = w
            out[f"{name}/{unc}_a"] = np.float64(params[unc]["a"])
            out[f"{name}/{unc}_b"] = np.float64(params[unc]["b"])
    return out


def task_cases(ref):
    """The reference's evaluation drivers (aggregate_uncertainties.py:133, prediction_shape_stats.py:70, ace.py:14/463,
    ncc.py:46, aurc.py:130) run unmodified on an in-memory experiment; only their third-party hooks (medpy's load,
    hydra's instantiate, jsbeautifier) are pointed at in-memory equivalents.  Records the JSON files they write."""
    import pathlib
    import types
    import warnings
    rng = np.random.default_rng(23)
    H, W, R, n_img = 24, 40, 3, 4
    ids = [f"img{i:02d}" for i in range(n_img)]
    data = {}
    for i in ids:
        mean_pred = (rng.random((H, W)) < 0.35).astype(np.uint8)
        mean_pred[8:16, 10:30] = 1
        preds = [np.where(rng.random((H, W)) < 0.9, mean_pred, 1 - mean_pred).astype(np.uint8) for _ in range(3)]
        maps = {u: (rng.random((H, W)) ** 3 * 0.69).astype(np.float32) for u in ("TU", "AU", "EU")}
        refs = np.stack([np.where(rng.random((H, W)) < np.clip(maps["TU"] * 2, 0.03, 0.8), 1 - mean_pred, mean_pred) for _ in range(R)]).astype(np.uint8)
        data[i] = dict(mean_pred=mean_pred, preds=preds, maps=maps, refs=refs)
    agg_cfg = {
        "patch_level": {"_target_": "evaluation.uncertainty_aggregation.aggregate_uncertainties.patch_level_aggregation", "patch_size": 10},
        "image_level": {"_target_": "evaluation.uncertainty_aggregation.aggregate_uncertainties.image_level_aggregation"},
        "threshold": {"_target_": "evaluation.uncertainty_aggregation.aggregate_uncertainties.threshold_aggregation"},
        "area_normalized": {"_target_": "evaluation.uncertainty_aggregation.aggregate_uncertainties.area_normalized_aggregation", "stats_filename": "area.json"},
        "border_normalized": {"_target_": "evaluation.uncertainty_aggregation.aggregate_uncertainties.border_normalized_aggregation", "stats_filename": "area.json"},
    }
    out = {"ids": np.array(ids), "agg_cfg": np.array(json.dumps(agg_cfg))}
    for i in ids:
        out[f"{i}/mean_pred"] = data[i]["mean_pred"]
        out[f"{i}/preds"] = np.stack(data[i]["preds"])
        out[f"{i}/refs"] = data[i]["refs"]
        for u in ("TU", "AU", "EU"):
            out[f"{i}/{u}"] = data[i]["maps"][u]
    with tempfile.TemporaryDirectory() as td:
        root = pathlib.Path(td)
        ds = root / "test"
        ds.mkdir()
        version = types.SimpleNamespace(unc_types=["TU", "AU", "EU"], exp_path=root, pred_model="Softmax", unc_ending=".tif",
                                        aggregations=["image_level", "threshold", "patch_level"], version_name="v0")
        loader = types.SimpleNamespace(
            exp_version=version, image_ids=ids, dataset_path=ds, unc_path_dict={u: ds / u for u in ("TU", "AU", "EU")},
            get_reference_segs=lambda i: data[i]["refs"], get_mean_pred_seg=lambda i: data[i]["mean_pred"],
            get_pred_segs=lambda i: data[i]["preds"], get_unc_map=lambda i, u: data[i]["maps"][u],
            get_gt_unc_map=lambda i: np.var(data[i]["refs"], axis=0), dataloader=None)
        thr = {"Softmax": {f"Mean {u} threshold": float(np.quantile(np.concatenate([data[i]["maps"][u].ravel() for i in ids]), 0.8)) for u in ("TU", "AU", "EU")}}
        json.dump(thr, open(root / "threshold_analysis.json", "w"))
        metrics = {i: {"metrics": {"dice": float(rng.random())}} for i in ids}
        json.dump(metrics, open(ds / "metrics.json", "w"))
        agg = ref.agg_module

        def fake_load(path):
            p = pathlib.Path(path)
            return data[p.name[: -len(".tif")]]["maps"][p.parent.name], None

        def fake_instantiate(cfg, **kw):
            cfg = dict(cfg)
            fn = getattr(agg, cfg.pop("_target_").rsplit(".", 1)[-1])
            return fn(**cfg, **kw)

        agg.load = fake_load
        agg.hydra.utils.instantiate = fake_instantiate
        agg.jsbeautifier.beautify = lambda s, opts: s
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref.shp_module.compute_prediction_shape_stats(loader)
            agg.aggregate_uncertainties(loader, agg_cfg)
            ref.ace_module.platt_scale_params(loader, ignore_value=None)
            ref.ace_module.calibration_error(loader, ignore_value=None)
            ref.ncc_module.main(loader)
            ref.aurc_module.main(loader)
        out["threshold_analysis.json"] = np.array(json.dumps(thr))
        out["metrics.json"] = np.array(json.dumps(metrics))
        for rel in ("test/area.json", "test/aggregated_TU.json", "test/aggregated_AU.json", "test/aggregated_EU.json",
                    "platt_scale_params.json", "test/calibration.json", "test/ambiguity_modeling.json", "test/failure_detection.json"):
            out[rel.split("/")[-1]] = np.array(open(root / rel).read())
    return out


def quantile_cases(ref):
    """eqACE (ace.py:378-406) on the per-(rater, pixel) arrays the body of calibration_error builds, and threshold
    discovery (find_threshold.py) through its own drivers on an in-memory experiment."""
    import pathlib
    import types
    rng = np.random.default_rng(29)
    out = {}
    specs = {
        "smooth": dict(H=24, W=32, R=4, n_cls=2, ignore=None, a=3.5, b=-1.25, kind="smooth"),
        "background_zeros": dict(H=32, W=32, R=4, n_cls=2, ignore=None, a=3.5, b=-1.25, kind="zeros"),  # 70 % u == 0: edges collapse
        "few_levels": dict(H=24, W=24, R=3, n_cls=2, ignore=None, a=5.0, b=-1.0, kind="levels"),     # 7 distinct values
        "gta_like_ignore": dict(H=24, W=32, R=5, n_cls=19, ignore=255, a=6.0, b=-2.0, kind="smooth"),
        "negative_a": dict(H=16, W=24, R=2, n_cls=2, ignore=None, a=-2.0, b=0.5, kind="smooth"),
        "steep": dict(H=16, W=24, R=2, n_cls=2, ignore=None, a=900.0, b=-300.0, kind="smooth"),
        "tiny": dict(H=1, W=3, R=1, n_cls=2, ignore=None, a=3.5, b=-1.25, kind="smooth"),               # fewer samples than bins
        "single": dict(H=1, W=1, R=1, n_cls=2, ignore=None, a=3.5, b=-1.25, kind="smooth"),
    }
    names = []
    for name, s in specs.items():
        H, W, R = s["H"], s["W"], s["R"]
        pred = rng.integers(0, s["n_cls"], (H, W)).astype(np.uint8)
        refs = np.stack([np.where(rng.random((H, W)) < 0.8, pred, rng.integers(0, max(s["n_cls"], 2), (H, W)))
                         for _ in range(R)]).astype(np.uint8)
        if s["ignore"] is not None:
            refs[rng.random(refs.shape) < 0.05] = s["ignore"]
        unc = (rng.random((H, W)) ** 2 * 0.69).astype(np.float32)
        if s["kind"] == "zeros":
            unc[rng.random((H, W)) < 0.7] = 0.0
        elif s["kind"] == "levels":
            unc = (np.round(unc * 9) / 9).astype(np.float32)
        with tempfile.TemporaryDirectory() as td:
            pf = os.path.join(td, "platt_scale_params.json")
            with open(pf, "w") as f:
                json.dump({"TU": {"a": s["a"], "b": s["b"]}}, f)
            pred_rep = np.repeat(pred[np.newaxis, :], R, 0)
            unc_rep = np.repeat(unc[np.newaxis, :], R, 0)
            correct = (refs == pred_rep).astype(int)
            if s["ignore"] is not None:
                keep = refs != s["ignore"]
                conf = ref.platt_scale_confid(-unc_rep[keep], platt_scale_file=pf, uncertainty="TU")
                cv = correct[keep]
            else:
                conf = ref.platt_scale_confid(-unc_rep.flatten(), platt_scale_file=pf, uncertainty="TU")
                cv = correct.flatten()
        names.append(name)
        out[f"{name}/refs"] = refs
        out[f"{name}/pred"] = pred
        out[f"{name}/unc"] = unc
        out[f"{name}/a"] = np.float64(s["a"])
        out[f"{name}/b"] = np.float64(s["b"])
        out[f"{name}/ignore"] = np.int64(-999 if s["ignore"] is None else s["ignore"])
        out[f"{name}/conf"] = conf
        out[f"{name}/correct"] = cv.astype(np.uint8)
        out[f"{name}/eqace"] = np.float64(ref.calc_eqace(cv, conf))
    out["eqace_cases"] = np.array(names)
    out["empty/eqace"] = np.float64(ref.calc_eqace(np.zeros(0, int), np.zeros(0, np.float32)))

    # np.quantile as find_threshold.py:76 calls it: float32 data, Python-float q
    maps = [(rng.random(shape) ** 3 * 0.69).astype(np.float32) for shape in ((24, 40), (17, 23), (8, 8, 8), (1, 5))]
    maps[1][rng.random(maps[1].shape) < 0.6] = 0.0
    qs = np.array([0.0, 1.0, 0.5, 0.25, 0.8, 0.937, 0.999, 1.0 / 3.0, 0.9183673469387755])
    for k, m in enumerate(maps):
        out[f"thr/map{k}"] = m
    out["thr/n_maps"] = np.int64(len(maps))
    out["thr/qs"] = qs
    with tempfile.TemporaryDirectory() as td:
        qf = os.path.join(td, "quantile_analysis.json")
        res_single, res_all = [], []
        for q in qs:
            with open(qf, "w") as f:
                json.dump({"Softmax": float(q)}, f)
            res_single.append([ref.calculate_threshold_image(qf, m, "Softmax") for m in maps])
            res_all.append(ref.calculate_threshold_image(qf, np.concatenate([m.ravel() for m in maps]), "Softmax"))
    out["thr/per_map"] = np.array(res_single, np.float64)
    out["thr/all_maps"] = np.array(res_all, np.float64)

    # the drivers: get_foreground_quantile -> save_foreground_quantiles -> threshold_images_paths -> find_threshold
    H, W, n_img = 24, 40, 5
    ids = [f"val{i:02d}" for i in range(n_img)]
    data = {}
    for i in ids:
        base = (rng.random((H, W)) < 0.2).astype(np.uint8)
        base[6:14, 8:24] = 1
        preds = [np.where(rng.random((H, W)) < 0.93, base, 1 - base).astype(np.uint8) * 255 for _ in range(3)]
        data[i] = dict(preds=preds, maps={u: (rng.random((H, W)) ** 3 * 0.69).astype(np.float32) for u in ("TU", "AU", "EU")})
        out[f"drv/{i}/preds"] = np.stack(preds)
        for u in ("TU", "AU", "EU"):
            out[f"drv/{i}/{u}"] = data[i]["maps"][u]
    out["drv/ids"] = np.array(ids)
    thr = ref.thr_module
    with tempfile.TemporaryDirectory() as td:
        root = pathlib.Path(td)
        ds = root / "val"
        ds.mkdir()
        version = types.SimpleNamespace(unc_types=["TU", "AU", "EU"], exp_path=root, pred_model="Softmax", unc_ending=".tif",
                                        version_name="v0")
        loader = types.SimpleNamespace(exp_version=version, image_ids=ids, dataset_path=ds,
                                       unc_path_dict={u: ds / u for u in ("TU", "AU", "EU")},
                                       get_pred_segs=lambda i: data[i]["preds"])

        def fake_load(path):
            p = pathlib.Path(path)
            return data[p.name[: -len(".tif")]]["maps"][p.parent.name], None

        thr.load = fake_load
        thr.save_foreground_quantiles(thr.get_foreground_quantile(loader))
        thr.find_threshold(thr.threshold_images_paths(loader))
        out["drv/quantile_analysis.json"] = np.array(open(root / "quantile_analysis.json").read())
        out["drv/threshold_analysis.json"] = np.array(open(root / "threshold_analysis.json").read())
    return out


def ged_nll_cases(ref):
    """ged_binary_fast (ged_fast.py:5-142) and Tester._compute_likelihood_stats / _compute_expected_nll
    (test_2D.py:1043-1120; unbound, with a stand-in ``self`` that only carries ignore_index)."""
    import types
    gen = torch.Generator().manual_seed(31)
    rng = np.random.default_rng(31)
    out = {}
    names = []
    am = ["dice", "max_dice_pred", "max_dice_gt", "major_dice"]
    specs = {
        "lidc_like": dict(P=8, C=2, G=4, S=(32, 32), ign=None, scale=2.0),
        "diffusion32": dict(P=32, C=2, G=4, S=(24, 24), ign=None, scale=1.0),
        "ignore255": dict(P=5, C=2, G=3, S=(16, 24), ign=255, scale=8.0),       # GED only: the reference's NLL cannot index 255
        "peaked": dict(P=6, C=2, G=2, S=(16, 16), ign=None, scale=60.0),        # p below the 1e-12 clamp
        "empty_pred": dict(P=4, C=2, G=3, S=(12, 12), ign=None, scale=2.0),     # both-empty / one-empty Dice rules
        "multiclass": dict(P=4, C=5, G=2, S=(12, 20), ign=None, scale=3.0),     # NLL only
        "single_rater_2d": dict(P=3, C=2, G=1, S=(9, 7), ign=None, scale=2.0),  # gt given as (H, W) to the likelihood functions
    }
    for name, sp in specs.items():
        P, C, G, S = sp["P"], sp["C"], sp["G"], sp["S"]
        logits = sp["scale"] * torch.randn(P, C, *S, generator=gen)
        if name == "empty_pred":
            logits[:, 1] -= 50.0  # every member predicts background everywhere
        x = torch.softmax(logits, dim=1)
        gt = torch.from_numpy(rng.integers(0, C, (G, *S))).long()
        if name == "empty_pred":
            gt[0] = 0  # one rater empty as well -> Dice 1 for that column, 0 for the others
        if sp["ign"] is not None:
            gt[torch.from_numpy(rng.random((G, *S)) < 0.1)] = sp["ign"]
        names.append(name)
        out[f"{name}/x"] = x.numpy()
        out[f"{name}/gt"] = gt.numpy()
        out[f"{name}/ignore"] = np.int64(-999 if sp["ign"] is None else sp["ign"])
        if C == 2:
            res = ref.ged_binary_fast(x, gt, ignore_index=sp["ign"], additional_metrics=am)
            for k in ["ged"] + am:
                out[f"{name}/{k}"] = np.float64(res[k])
        if sp["ign"] is None:
            stand_in = types.SimpleNamespace(ignore_index=-1)
            gt_arg = gt[0] if name == "single_rater_2d" else gt
            m, g, mean = ref.Tester._compute_likelihood_stats(stand_in, x, gt_arg)
            out[f"{name}/gt_model_nll"] = np.array(m, np.float64)
            out[f"{name}/gt_nll"] = np.array(g, np.float64)
            out[f"{name}/mean_nll"] = np.float64(mean)
            out[f"{name}/expected_nll"] = np.float64(ref.Tester._compute_expected_nll(stand_in, x, gt_arg))
    # an ignore value that IS a valid class index (the only way the reference's NLL runs with ignore_index >= 0)
    P, C, G, S = 4, 3, 2, (10, 14)
    x = torch.softmax(2.0 * torch.randn(P, C, *S, generator=gen), dim=1)
    gt = torch.from_numpy(rng.integers(0, C, (G, *S))).long()
    gt[1][:] = 2  # rater 1 entirely ignored -> valid_count == 0 -> zeros (test_2D.py:1060-1061)
    stand_in = types.SimpleNamespace(ignore_index=2)
    m, g, mean = ref.Tester._compute_likelihood_stats(stand_in, x, gt)
    names.append("ignore_is_class")
    out["ignore_is_class/x"] = x.numpy()
    out["ignore_is_class/gt"] = gt.numpy()
    out["ignore_is_class/ignore"] = np.int64(2)
    out["ignore_is_class/gt_model_nll"] = np.array(m, np.float64)
    out["ignore_is_class/gt_nll"] = np.array(g, np.float64)
    out["ignore_is_class/mean_nll"] = np.float64(mean)
    out["ignore_is_class/expected_nll"] = np.float64(ref.Tester._compute_expected_nll(stand_in, x, gt))
    out["cases"] = np.array(names)
    return out


def main():
    assert ref_shim.available(), "needs /root/reference"
    ref = ref_shim.load()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(1)  # thread-count independent reduction rows (see oracle.cascade_sum_f32)
    only = set(sys.argv[1:])  # e.g. ``python -m oracle.make_golden quantile.npz``: regenerate one file
    for fname, builder in (("uncertainty.npz", uncertainty_cases), ("aggregation.npz", aggregation_cases),
                           ("calibration.npz", calibration_cases), ("ncc_aurc.npz", ncc_aurc_cases),
                           ("platt_fit.npz", platt_fit_cases), ("tasks.npz", task_cases), ("quantile.npz", quantile_cases), ("ged_nll.npz", ged_nll_cases)):
        if only and fname not in only:
            continue
        data = builder(ref)
        np.savez_compressed(os.path.join(GOLDEN, fname), **data)
        print(fname, len(data), "arrays", os.path.getsize(os.path.join(GOLDEN, fname)), "bytes")
    import scipy
    import sklearn
    manifest = {
        "generator": "oracle/make_golden.py",
        "reference": "JakobLC/DiffUncertainty @ /root/reference (unmodified functions via oracle/ref_shim.py)",
        "torch": torch.__version__, "numpy": np.__version__, "scipy": scipy.__version__,
        "sklearn": sklearn.__version__, "torch_threads": torch.get_num_threads(),
    }
    with open(os.path.join(GOLDEN, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
