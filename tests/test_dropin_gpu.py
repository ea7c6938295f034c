"""GPU: the drop-in boundary exercised by the reference's OWN callers.  With `shim/` in front of the staged reference
(baseline/_ref, tools/stage_reference.py) on PYTHONPATH, `from model import build_segmenter` resolves to
cris.pytorch_b200 while `utils.config`, `engine.engine.train` and the yaml files are the reference's, unmodified:
the test loads config/refcoco/cris_r50.yaml with the reference's loader, builds the model through the reference's
entry point, and runs two iterations of the reference's `engine.train` (SyncBN conversion + DDP + Adam + MultiStepLR
+ GradScaler + amp.autocast, train.py:94-111, engine/engine.py:17-88) and the tools/latency.py loop on it."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(REPO, "baseline", "_ref")


def test_reference_train_loop_and_latency_tool_run_on_the_dropin():
    if not os.path.isfile(os.path.join(REF, "engine", "engine.py")):
        pytest.skip("baseline/_ref is not staged (python tools/stage_reference.py in the build container)")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(REPO, "shim"), REPO, REF, os.path.join(REPO, "tests", "stubs")])
    env["CRIS_REF_ROOT"] = REF
    env["WANDB_MODE"] = "disabled"
    r = subprocess.run([sys.executable, os.path.join(REPO, "tests", "dropin_driver.py"), "r50", "4"], env=env,
                       capture_output=True, text=True, timeout=900, cwd=REF)
    line = next((ln for ln in r.stdout.splitlines() if ln.startswith("DROPIN ")), None)
    assert r.returncode == 0 and line is not None, (r.stdout[-2000:], r.stderr[-4000:])
    out = json.loads(line[len("DROPIN "):])
    assert out["train_iters"] == 2 and out["params_changed"] == 4, out
    assert out["pred_shape"] == [1, 1, 104, 104]
    assert 140 < out["n_params_M"] < 150      # cris_r50: 146.85 M trainable parameters
    assert out["latency_b1_ms"] < 50


def test_staged_reference_is_byte_identical_to_its_manifest():
    if not os.path.isfile(os.path.join(REF, "MANIFEST.json")):
        pytest.skip("baseline/_ref is not staged")
    import hashlib
    man = json.load(open(os.path.join(REF, "MANIFEST.json")))["files"]
    assert len(man) >= 20
    for rel, sha in man.items():
        assert hashlib.sha256(open(os.path.join(REF, rel), "rb").read()).hexdigest() == sha, rel
