"""The evaluation-task mirrors (diffuncertainty_b200.tasks) against the JSON files the reference's own drivers wrote
for the same in-memory experiment (tests/golden/tasks.npz, recorded by oracle/make_golden.py::task_cases)."""
import json
import os
import pathlib
import tempfile
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu


def assert_json_close(got, want, path="", rtol=1e-5, skip=(), atol=1e-12):
    if isinstance(want, dict):
        assert isinstance(got, dict) and set(got) == set(want), (path, sorted(got), sorted(want))
        for k in want:
            if k in skip:
                continue
            assert_json_close(got[k], want[k], f"{path}/{k}", rtol, skip, atol)
    elif isinstance(want, (list, tuple)):
        assert len(got) == len(want), path
        for i, (g, w) in enumerate(zip(got, want)):
            assert_json_close(g, w, f"{path}[{i}]", rtol, skip, atol)
    elif isinstance(want, float):
        np.testing.assert_allclose(got, want, rtol=rtol, atol=atol, err_msg=path)
    else:
        assert got == want, (path, got, want)


@pytest.fixture(scope="module")
def experiment():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    with np.load(os.path.join(GOLDEN_DIR, "tasks.npz")) as z:
        g = {k: z[k] for k in z.files}
    ids = [str(i) for i in g["ids"]]
    td = tempfile.TemporaryDirectory()
    root = pathlib.Path(td.name)
    ds = root / "test"
    ds.mkdir()
    (root / "threshold_analysis.json").write_text(str(g["threshold_analysis.json"]))
    (ds / "metrics.json").write_text(str(g["metrics.json"]))
    version = types.SimpleNamespace(unc_types=["TU", "AU", "EU"], exp_path=root, pred_model="Softmax", unc_ending=".tif",
                                    aggregations=["image_level", "threshold", "patch_level"], version_name="v0")
    loader = types.SimpleNamespace(
        exp_version=version, image_ids=ids, dataset_path=ds, unc_path_dict={u: ds / u for u in ("TU", "AU", "EU")},
        get_reference_segs=lambda i: g[f"{i}/refs"], get_mean_pred_seg=lambda i: g[f"{i}/mean_pred"],
        get_pred_segs=lambda i: list(g[f"{i}/preds"]), get_unc_map=lambda i, u: g[f"{i}/{u}"],
        get_gt_unc_map=lambda i: np.var(g[f"{i}/refs"], axis=0), load_unc_file=lambda u, i: g[f"{i}/{u}"], dataloader=None)
    yield g, loader, root, ds
    td.cleanup()


def test_shape_stats_and_aggregation_tasks(experiment):
    from diffuncertainty_b200 import tasks
    g, loader, root, ds = experiment
    stats = tasks.compute_prediction_shape_stats(loader)
    assert_json_close(json.loads((ds / "area.json").read_text()), json.loads(str(g["area.json"])))
    assert stats == json.loads((ds / "area.json").read_text())
    tasks.aggregate_uncertainties(loader, json.loads(str(g["agg_cfg"])))
    for unc in ("TU", "AU", "EU"):
        got = json.loads((ds / f"aggregated_{unc}.json").read_text())
        want = json.loads(str(g[f"aggregated_{unc}.json"]))
        assert_json_close(got, want, unc)   # scores within 1e-5, bounding boxes and thresholds exact


def test_platt_and_calibration_tasks(experiment):
    from diffuncertainty_b200 import calibration, tasks
    g, loader, root, ds = experiment
    params = calibration.platt_scale_params(loader, ignore_value=None)
    want = json.loads(str(g["platt_scale_params.json"]))
    assert_json_close(params, want, rtol=2e-4)
    (root / "platt_scale_params.json").write_text(str(g["platt_scale_params.json"]))  # evaluate with the reference's own fit
    tasks.calibration_error(loader, ignore_value=None)
    # calibration errors are differences of two means in [0, 1]: compare them on that scale (the AU fit of this
    # experiment saturates: every sample sits in the last bin and |acc - conf| is ~1e-8)
    assert_json_close(json.loads((ds / "calibration.json").read_text()), json.loads(str(g["calibration.json"])), rtol=1e-5, atol=1e-6)


def test_ncc_and_aurc_tasks(experiment):
    from diffuncertainty_b200 import tasks
    g, loader, root, ds = experiment
    tasks.ncc_main(loader)
    assert_json_close(json.loads((ds / "ambiguity_modeling.json").read_text()), json.loads(str(g["ambiguity_modeling.json"])))
    tasks.aggregate_uncertainties(loader, json.loads(str(g["agg_cfg"])))
    tasks.aurc_main(loader)
    assert_json_close(json.loads((ds / "failure_detection.json").read_text()), json.loads(str(g["failure_detection.json"])), rtol=1e-4)
