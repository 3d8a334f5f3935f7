"""csrc/std_sort_replay.cuh -- the code the device runs to reproduce the reference's two std::sort calls
(src/match/match_features.cpp:100-101, src/model_inliers/ransac.cpp:83-90) -- compiled for the HOST and checked against
libstdc++'s std::sort itself: tests/sort_replay_check.cpp (random keys with ties, sorted, reversed, organ pipe, all
equal, and McIlroy-adversary inputs that force the heap-sort fallback)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_replay_equals_libstdcxx_sort(tmp_path):
    exe = str(tmp_path / "sort_replay_check")
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror",
                        os.path.join(ROOT, "tests", "sort_replay_check.cpp"), "-o", exe],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout[-2000:]
    assert "heap-sort fallback taken" in r.stdout
