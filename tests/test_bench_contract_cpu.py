"""bench.py's reference arm runs on host cores only, so its output contract can be checked without a GPU."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                          "--ref-batch", "2"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("image-query-pairs/sec") and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert "workload" in d["config"] and "model" not in d["config"]
    # the reference arm describes the SAME workload as the zsg arm (same config dict), and says what it sampled of it
    sys.path.insert(0, ROOT)
    import bench
    args = argparse.Namespace(model="retina", batch=64, dtype="fp32")
    assert d["config"] == bench.config_dict(args, 1)
    cb = d["cpu_baseline"]
    from oracle import ref_harness
    want_kind = "reference" if ref_harness.find_reference() is not None else "port"
    assert cb["kind"] == want_kind and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert d["sample_batch"] == 2 and "each step = 2 pairs" in cb["sample"] and "bs=64" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0 and not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]


def test_reference_tree_copy_is_kept_out_of_history():
    """oracle/_ref (the copy of the reference's sources the reference arm runs on the GPU box) is git-ignored but travels
    with gpurun snapshots."""
    gi = open(os.path.join(ROOT, ".gitignore")).read().split()
    gr = open(os.path.join(ROOT, ".gpurunignore")).read().split()
    assert "oracle/_ref/" in gi and not any(x.startswith("oracle") for x in gr)
