"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`, the CPU restatement of obake's
poly_mul_impl_mt_hm on the host cores) prints ONE JSON line with the keys the driver reads, and the algorithmic-bytes
model matches SURVEY 8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--workload", "F20"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "term_products_per_s" and d["unit"] == "term-products/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["sample"] == "full workload"  # the reference arm runs the same configuration, never a sample
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "F20" and "model" not in d["config"]
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--workload", "F20"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_algorithmic_bytes_model():
    sys.path.insert(0, ROOT)
    import bench

    # B = K_in + C_in + K_slot + 2*A
    assert bench.algorithmic_bytes(1, False, 1, 2) == 56   # F20, S12, S16T
    assert bench.algorithmic_bytes(1, False, 1, 3) == 72   # F30
    assert bench.algorithmic_bytes(1, False, 1, 1) == 40   # dense_02
    assert bench.algorithmic_bytes(1, True, 1, 1) == 40    # fp64, one packed word
    assert bench.algorithmic_bytes(2, True, 1, 1) == 56    # fp64, d_packed with two words


def test_multi_gpu_also_deadline_keeps_the_headline_line():
    """bench.also_multi_gpu: if the extra workloads hang (a peer stuck in a collective), the deadline prints the
    headline line that is already complete -- ONE JSON line, exit code 0 -- instead of losing it."""
    code = r"""
import json, sys, time, types
sys.path.insert(0, %r)
import bench
def hang(*a, **k):
    time.sleep(3600)
bench.bench_workload = hang
args = types.SimpleNamespace(also="S12,RECT", workload="F30", also_deadline=1.0, steps=4)
line = {"metric": bench.METRIC, "value": 1.0, "n_gpus": 2}
bench.also_multi_gpu(args, None, None, None, 0, 2, line)
print("NOT REACHED")
""" % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.strip()]
    assert len(lines) == 1 and "NOT REACHED" not in r.stdout
    d = json.loads(lines[0])
    assert d["value"] == 1.0 and d["n_gpus"] == 2 and "deadline" in d["also"]["error"]
    # a rank other than 0 leaves quietly
    r = subprocess.run([sys.executable, "-c", code.replace("None, 0, 2, line", "None, 1, 2, line")], capture_output=True,
                       text=True, timeout=60, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_multi_gpu_also_failure_on_one_rank_keeps_the_headline_line():
    code = r"""
import json, sys, types
sys.path.insert(0, %r)
import bench
def boom(*a, **k):
    raise RuntimeError("segment tables keep overflowing")
bench.bench_workload = boom
args = types.SimpleNamespace(also="S12", workload="F30", also_deadline=30.0, steps=4)
bench.also_multi_gpu(args, None, None, None, 0, 2, {"metric": bench.METRIC, "value": 2.0})
""" % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["value"] == 2.0 and "RuntimeError" in d["also"]["error"]
