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
