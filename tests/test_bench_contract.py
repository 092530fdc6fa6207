"""bench.py's reference arm (the CPU leg the driver launches next to ours) prints ONE JSON line with the contract's
keys; under a multi-rank launch only rank 0 works and prints.  CPU only (two steps of the oracle port, a few seconds)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "1"], capture_output=True, text=True, env=env, timeout=300, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    r = _run({})
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["metric"] == "ppo_samples_per_s_fwd_bptt" and d["unit"] == "samples/s" and d["value"] > 0
    assert d["config"]["workload"] == "a1_amp_ppo_update_4096x24"
    from oracle import ref_loader as R
    # the UNMODIFIED reference when oracle/make_ref.sh staged it (build container and GPU box), else the oracle port
    assert d["cpu_baseline"]["kind"] == ("reference" if R.available("amp") else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["fwd"]["value"] > 0 and d["fwd"]["kind"] == ("reference" if R.available("rma") else "port")
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
