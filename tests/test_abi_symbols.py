"""The C-ABI library loads without a GPU and exports every symbol include/ssam_gpu.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

from solidstateamfoam_b200 import capi
from solidstateamfoam_b200 import params as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ssam_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"SSAM_API\s+[\w\s\*]+?\b(ssam_\w+)\s*\(", src)))


def test_header_symbols_are_exported():
    names = declared_symbols()
    assert len(names) >= 30
    L = capi.load()
    for n in names:
        assert hasattr(L, n), f"{n} declared in ssam_gpu.h but not exported"
    assert sorted(capi.SYMBOLS) == names, "capi.SYMBOLS out of date with the header"
    assert L.ssam_abi_version() == 1


def test_no_extra_dynamic_dependencies():
    """cudart is linked statically and NCCL is dlopen'ed, so the library loads on a CPU-only host."""
    out = subprocess.run(["ldd", capi.lib_path()], capture_output=True, text=True).stdout
    assert "libcudart" not in out and "libnccl" not in out and "libtorch" not in out


def test_struct_layouts_match_header():
    """sizes the C compiler gives the POD structs == the ctypes mirrors (catches field drift)."""
    code = r'''
#include <stdio.h>
#include "ssam_gpu.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(ssam_mesh_desc), sizeof(ssam_viscosity_params),
         sizeof(ssam_conductivity), sizeof(ssam_mixture_params), sizeof(ssam_stick_params),
         sizeof(ssam_rotslip_params), sizeof(ssam_update_params), sizeof(ssam_host_io));
  return 0; }
'''
    d = os.path.join(ROOT, ".pytest_cache")
    os.makedirs(d, exist_ok=True)
    c = os.path.join(d, "layout.c")
    exe = os.path.join(d, "layout")
    open(c, "w").write(code)
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, c], check=True)
    sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    mirrors = [P.MeshDesc, P.ViscosityParams, P.Conductivity, P.MixtureParams, P.StickParams, P.RotSlipParams,
               P.UpdateParams, P.HostIO]
    assert sizes == [ctypes.sizeof(m) for m in mirrors]


def test_create_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.SsamError) as e:
        capi.Context(0)
    assert e.value.status == P.ERR_CUDA and "no CPU fallback" in str(e.value)
