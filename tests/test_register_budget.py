"""Register and stack budgets of the shipped kernels, read from the library with cuobjdump (no GPU needed).

Occupancy is part of the measured performance of these kernels: the pair kernel is tuned for 128 registers = 4 CTAs
of 128 threads per SM, the Euler projection kernels for 128 registers = 3 CTAs of 160 threads, the PhysicalOperators kernels of
BASELINE config 3 for 40 registers = 48 warps per SM.  A source change that silently moves one of them to another register regime
(as the two-partial-sum form of the wide rows did to the unbounded PhysicalOperators kernels: 0.86 -> 1.18 ms on config 3) shows
up here on the CPU, before anything is timed."""
import re
import shutil
import subprocess

import pytest

from sse_b200 import _lib


def resources():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        out = subprocess.run([exe, "--dump-resource-usage", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    except (OSError, subprocess.CalledProcessError):
        pytest.skip("cuobjdump not available")
    table = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", out):
        table[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    assert table, "no kernels found in " + _lib.LIB_PATH
    return table


def pick(table, pattern):
    hits = {k: v for k, v in table.items() if re.search(pattern, k)}
    assert hits, pattern
    return hits


def test_headline_kernels_keep_their_register_regime():
    t = resources()
    for name, (reg, stack) in pick(t, r"k_fluxdiff_ctILi5ELi4ELb[01]E").items():       # pair kernel, N = 5 (both variants)
        assert reg <= 128 and stack == 0, (name, reg, stack)
    for name, (reg, stack) in pick(t, r"k_nodal_ctILi5ELi5ELi3ELb1ELb[01]E").items():   # pass A and the stage-fused form
        assert reg <= 128 and stack <= 32, (name, reg, stack)
    for name, (reg, stack) in pick(t, r"k_project_ctILi5ELi5ELi3E").items():
        assert reg <= 128 and stack == 0, (name, reg, stack)


def test_physical_operator_kernels_of_config3_stay_at_48_warps_per_sm():
    t = resources()
    for name, (reg, stack) in pick(t, r"k_(aux|time)_physicalILi2ELi1E").items():
        assert reg <= 40 and stack <= 64, (name, reg, stack)


def test_every_compile_time_degree_is_instantiated():
    t = resources()
    for n in range(3, 9):                                                               # N = p + 1 = 3 .. 8
        pick(t, rf"k_fluxdiff_ctILi{n}E")
        pick(t, rf"k_nodal_ctILi{n}ELi5E")
        pick(t, rf"k_project_ctILi{n}ELi5E")
        pick(t, rf"k_adv_fused_ctILi{n}E")
        pick(t, rf"k_adv_facets_ctILi{n}E")
    for n in range(3, 6):                                                               # triangles: N = 3 .. 5
        pick(t, rf"k_tri_fluxdiffILi{n}E")
        pick(t, rf"k_tri_advILi{n}E")
    # the pair kernel of the degrees added last: no spills at N = 6, 7 (two CTAs per SM)
    for name, (reg, stack) in pick(t, r"k_fluxdiff_ctILi[67]ELi2ELb0E").items():
        assert reg <= 128, (name, reg, stack)
