"""bench.py's reference arm (`--impl reference`) on a tiny workload, on the CPU: one JSON line with the keys the driver
reads, no product library mapped into the process (the arm times the CPU restatement of the reference alone), and the
silent exit of ranks other than 0 under torchrun."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--channels", "8", "--seconds", "0.25", "--steps", "2",
                           "--warmup", "1"], capture_output=True, text=True, cwd=ROOT, env=env, timeout=300)


def test_reference_arm_prints_one_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "GC-ADPCM encode Msamples/sec (batch)" and line["unit"] == "Msamples/s"
    assert line["higher_is_better"] is True and line["steps"] == 2 and line["warmup"] == 1 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_reference_arm_other_ranks_exit_silently():
    r = _run({"RANK": "3", "LOCAL_RANK": "3", "WORLD_SIZE": "8"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_does_not_map_the_product_library():
    code = ("import sys, runpy\n"
            "sys.argv = ['bench.py', '--impl', 'reference', '--channels', '4', '--seconds', '0.1', '--steps', '1', '--warmup', '0']\n"
            "try:\n    runpy.run_path('bench.py', run_name='__main__')\nexcept SystemExit:\n    pass\n"
            "maps = open('/proc/self/maps').read()\n"
            "print('MAPPED' if 'libvgaudio_b200' in maps else 'CLEAN')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip().endswith("CLEAN")
