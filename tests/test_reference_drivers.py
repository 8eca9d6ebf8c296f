"""The reference's drivers, UNCHANGED (build container only: they are read from /root/reference).

spmm_test.py and inference.py are launched through tests/shims/run_reference.py, which only prepares the imports
(stand-ins for torch_sparse / torch_geometric / ogb, `models` as a package).  `--version cpu` is the reference's own
CPU path (`torch_sparse.matmul`, served by the oracle).  Their stdout is parsed with the reference's own
Experiment.parse_result (utils/experiment.py:468-491) - the `[DATA]key: value` protocol SURVEY.md 8f-4 asks to keep -
and so is the stdout our GPU driver examples/spmm_test.py produced on the B200 box (tests/golden/spmm_test_gpu_stdout.txt)."""
import os
import subprocess
import sys
import types

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("PYGIM_REFERENCE_ROOT", "/root/reference")
needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "spmm_test.py")), reason="needs the reference tree")


def _run(script, *args, cwd):
    cmd = [sys.executable, os.path.join(HERE, "shims", "run_reference.py"), os.path.join(REF, script)] + list(args)
    env = dict(os.environ, PYGIM_SHIM_SCALE="0.05")
    r = subprocess.run(cmd, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    return r.stdout


def _parse_with_reference(stdout_text, tmp_path):
    sys.path.insert(0, os.path.join(HERE, "shims"))
    import run_reference
    run_reference.prepare()
    from utils.experiment import Experiment                     # the reference's own parser
    path = tmp_path / "run.out"
    path.write_text(stdout_text)
    stub = types.SimpleNamespace(stdout_path=lambda result_root: str(path))
    return Experiment.parse_result(stub, str(tmp_path))


@needs_ref
def test_reference_spmm_test_cpu_version_runs_unchanged(tmp_path):
    out = _run("spmm_test.py", "--version", "cpu", "--dataset", "PubMed", "--data_type", "FLT32", "--hidden_size", "32",
               "--repeat", "2", cwd=str(tmp_path))
    assert out.count("[DATA]torch_time(ms)") == 2 and "Model=spmm_test Repeat=1" in out
    res = _parse_with_reference(out, tmp_path)
    assert res["repeat"] == 2 and res["torch_time(ms)"] > 0


@needs_ref
@pytest.mark.parametrize("model", ["gcn", "gin", "sage"])
def test_reference_inference_cpu_version_runs_unchanged(tmp_path, model):
    out = _run("inference.py", "--version", "cpu", "--dataset", "PubMed", "--model", model, "--hidden_size", "32",
               "--data_type", "INT32", "--repeat", "2", cwd=str(tmp_path))
    assert out.count("[DATA]infer_time(ms)") == 2 and "Test_acc" in out
    res = _parse_with_reference(out, tmp_path)
    assert res["repeat"] == 2 and res["infer_time(ms)"] > 0


@needs_ref
def test_reference_parser_reads_our_gpu_drivers_stdout(tmp_path):
    """examples/spmm_test.py keeps the reference's stdout protocol; its B200 output (captured on the GPU box by
    tools/gpu_r2_*.sh) parses with the reference's Experiment.parse_result, phase timers included."""
    golden = os.path.join(HERE, "golden", "spmm_test_gpu_stdout.txt")
    if not os.path.exists(golden):
        pytest.skip("no captured GPU driver output yet")
    res = _parse_with_reference(open(golden).read(), tmp_path)
    assert res["repeat"] >= 2
    for key in ("pim_time_spmm(ms)", "kernel_time", "load_dense_time", "retrieve_result_time"):
        assert key in res and res[key] >= 0, (key, sorted(res))
