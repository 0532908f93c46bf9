"""Host-side pieces that need no GPU: the NUMA binding helper degrades to a no-op, and bench.py's reference arm (the oracle
port on the host cores, the one place outside tests/ that may execute oracle/) prints one well-formed JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_hostbind_is_a_noop_without_a_gpu():
    from qa_tiger_b200 import hostbind
    before = sorted(os.sched_getaffinity(0))
    assert hostbind.gpu_local_cpus(0) is None or isinstance(hostbind.gpu_local_cpus(0), list)
    prev = hostbind.bind_to_gpu(0)
    after = sorted(os.sched_getaffinity(0))
    if prev is None:
        assert after == before                  # nothing known about the device: affinity untouched
    else:
        assert set(after) <= set(before)
        os.sched_setaffinity(0, prev)


def test_hostbind_can_be_switched_off(monkeypatch):
    from qa_tiger_b200 import hostbind
    monkeypatch.setenv("QAT_HOSTBIND", "0")
    assert hostbind.bind_to_gpu(0) is None


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "questions/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["gpu_launches"] == 0


def test_tome_frame_shards_are_contiguous_and_cover_every_frame():
    """configs[2] shards by frame, no exchange -- rank r owns [r*ceil(N/W), (r+1)*ceil(N/W)), in order, no overlap, no
    padding; ranks beyond the frames own nothing."""
    from qa_tiger_b200.tome.merge import frame_shard
    for n in (0, 1, 7, 592, 15360, 15361):
        for world in (1, 2, 3, 8, 16):
            shards = [frame_shard(n, r, world) for r in range(world)]
            covered = [f for lo, hi in shards for f in range(lo, hi)]
            assert covered == list(range(n)), (n, world)
            sizes = [hi - lo for lo, hi in shards]
            assert max(sizes) - min(s for s in sizes if s or n < world) <= -(-n // world)
            assert all(lo <= hi for lo, hi in shards)
    assert frame_shard(15360, 3, 8) == (5760, 7680)
