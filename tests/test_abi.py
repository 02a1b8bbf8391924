"""CPU: the C-ABI library loads and exports every symbol include/pinnjet.h declares (no compute calls without a GPU);
the ctypes mirror of PjSpec has the size the C compiler gives the struct."""
import ctypes
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    sys.path.insert(0, ROOT)
    from neurodiffeq_b200.csrc import build as pj_build
    lib_path = pj_build.build()
    header = open(os.path.join(ROOT, "include", "pinnjet.h")).read()
    declared = set(re.findall(r"\b(pj_[a-z_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(lib_path)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} not exported"
    from neurodiffeq_b200 import engine
    assert set(engine.EXPORTED_SYMBOLS) == declared
    assert lib.pj_abi_version() == 2


def test_ctypes_struct_layout_matches_c(tmp_path):
    from neurodiffeq_b200 import engine
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "pinnjet.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", '
                   'sizeof(PjNet), sizeof(PjSpec), sizeof(PjSizes), offsetof(PjSpec, dir), offsetof(PjSpec, n_theta), '
                   'offsetof(PjSpec, net), offsetof(PjNet, w_off), offsetof(PjNet, yrow0));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    a, b, c, o_dir, o_theta, o_net, o_woff, o_yrow = map(int, subprocess.check_output([str(exe)]).split())
    assert (a, b, c) == (ctypes.sizeof(engine.PjNet), ctypes.sizeof(engine.PjSpec), ctypes.sizeof(engine.PjSizes))
    assert (o_dir, o_theta, o_net) == (engine.PjSpec.dir.offset, engine.PjSpec.n_theta.offset, engine.PjSpec.net.offset)
    assert (o_woff, o_yrow) == (engine.PjNet.w_off.offset, engine.PjNet.yrow0.offset)


def test_sampler_struct_layout_matches_c(tmp_path):
    from neurodiffeq_b200 import device_sampling as ds
    src = tmp_path / "sz2.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "pinnjet.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", '
                   'sizeof(PjSampleLaw), sizeof(PjSampler), offsetof(PjSampleLaw, div), offsetof(PjSampleLaw, base), '
                   'offsetof(PjSampler, law));return 0;}\n')
    exe = tmp_path / "sz2"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    a, b, o_p0, o_base, o_law = map(int, subprocess.check_output([str(exe)]).split())
    assert (a, b) == (ctypes.sizeof(ds.PjSampleLaw), ctypes.sizeof(ds.PjSampler))
    assert (o_p0, o_base, o_law) == (ds.PjSampleLaw.div.offset, ds.PjSampleLaw.base.offset, ds.PjSampler.law.offset)


def test_no_gpu_means_loud_failure():
    import torch
    import pytest
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from helpers import build_fused
    with pytest.raises(RuntimeError, match="CUDA device"):
        build_fused("c2")
