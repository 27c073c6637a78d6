"""CPU: the C-ABI library loads, exports every symbol include/mds_b200.h declares, and the ctypes
mirrors of its structs have the C compiler's sizes.  No compute calls (no GPU here)."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mds_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mds_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib_built):
    from multidronesim_b200 import _lib
    names = declared_functions()
    assert len(names) >= 27
    for n in names:
        assert hasattr(lib_built, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == names  # the ctypes table covers exactly the header
    assert lib_built.mds_abi_version() == 9
    assert lib_built.mds_cbf_num_rows(2, 8, 1) == 100 and lib_built.mds_cbf_num_rows(3, 8, 1) == 116  # SURVEY App. C row counts
    assert lib_built.mds_cbf_num_rows(2, 2, 1) == 19 and lib_built.mds_cbf_num_rows(3, 7, 0) == 91


def test_struct_sizes_match_c():
    from multidronesim_b200 import _lib
    prog = r'''
#include <stdio.h>
#include "mds_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(MdsRlsCfg), sizeof(MdsDslPidState), sizeof(MdsDslPidGains), sizeof(MdsDroneParams), sizeof(MdsState), sizeof(MdsPidState),
         sizeof(MdsGeoGains), sizeof(MdsLqrGains), sizeof(MdsCbfParams), sizeof(MdsRolloutCfg), sizeof(MdsTrajSpecF32),
         sizeof(MdsTrajSpecF64), sizeof(MdsTrajSegF32), sizeof(MdsTrajSegF64));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)  # header is plain C
        sizes = [int(x) for x in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()]
    py = [ctypes.sizeof(t) for t in (_lib.RlsCfg, _lib.DslPidState, _lib.DslPidGains, _lib.DroneParams, _lib.State, _lib.PidState, _lib.GeoGains, _lib.LqrGains, _lib.CbfParams, _lib.RolloutCfg)]
    py += [_lib.traj_spec_dtype("f4").itemsize, _lib.traj_spec_dtype("f8").itemsize, _lib.traj_seg_dtype("f4").itemsize, _lib.traj_seg_dtype("f8").itemsize]
    assert py == sizes


def test_no_cpu_fallback(lib_built):
    """Product code must fail loudly without a GPU and must never import the oracle."""
    import torch
    import multidronesim_b200 as mds
    if not torch.cuda.is_available():
        with pytest.raises(mds._lib.MdsError):
            mds.BatchedCtrlAviary(num_drones=1)
    pkg = os.path.join(ROOT, "multidronesim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"


def test_argument_errors_do_not_need_a_gpu(lib_built):
    """status codes / messages of the boundary (no launch happens on a NULL-pointer call)"""
    from multidronesim_b200 import _lib
    from multidronesim_b200.constants import DroneConstants
    prm = DroneConstants().c_params()
    rc = lib_built.mds_physics_step_f32(prm, _lib.State(None, None, None, None, None), None, None, None, 4, 2, None)
    assert rc == -1 and b"null pointer" in lib_built.mds_last_error()
    c = _lib.CbfParams()
    c.order = 5
    rc = lib_built.mds_cbf_rows_f64(prm, c, None, None, None, 0, None, None, 1, 2, None)
    assert rc == -1 and b"order" in lib_built.mds_last_error()
