"""The 32-lane form of the striped local SW (two warp lanes per striped segment, ssq_warp.cuh: sw_local_pass_warp_split) is device-only
code; its algorithm — the repair of a segment's second half by the first half's F and the two-step lazy-F sweep — is restated lane by
lane in tests/hostsim/split_emul.cpp and checked here against the scalar restatement of the 16-lane kernel (sw_local_pass,
ssq_dev2.cuh, itself equal to the oracle's ksw_align2 on the GPU: tests/test_gpu_parity.py::test_sw_local_striped_order)."""
import os
import subprocess

import ssq_testlib as T


def test_half_segment_form_equals_the_16_lane_kernel(tmp_path):
    exe = str(tmp_path / "split_emul")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(T.ROOT, "include"), "-o", exe, os.path.join(T.ROOT, "tests", "hostsim", "split_emul.cpp")], check=True)
    p = subprocess.run([exe, "12000"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert p.returncode == 0, p.stderr.decode()
    out = p.stdout.decode()
    assert " 0 mismatches" in out
    # the random problems must reach every path: repairs of second halves, lazy sweeps past the first cell, into the second halves, later rounds
    nums = [int(x) for x in out.replace(";", " ").replace(",", " ").split() if x.isdigit()]
    assert all(n > 1000 for n in nums[4:8]), out
