"""The C-ABI boundary without a GPU: libmobi_b200.so loads, exports every function include/mobi_b200.h declares, the
ctypes binding (mobi_b200/_lib.py) names exactly that set, and the ctypes structures have the size the C compiler gives
the header's structs (checked by compiling a tiny sizeof program with gcc).  No compute call is made here."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mobi_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mobi_[a-z0-9_]+)\s*\(", src)))


def _lib():
    from mobi_b200 import _lib, build
    build.build()  # no-op when the library is current; compiles with nvcc otherwise (no GPU needed)
    return _lib, _lib.load()


def test_library_exports_every_declared_symbol():
    L, lib = _lib()
    declared = _declared_functions()
    assert len(declared) >= 38
    for name in declared:
        assert hasattr(lib, name), "libmobi_b200.so does not export %s" % name
    assert sorted(L.SIGNATURES) == declared, (sorted(set(declared) ^ set(L.SIGNATURES)))
    assert lib.mobi_version() >= 100
    assert isinstance(lib.mobi_last_error(), bytes)


def test_ctypes_struct_sizes_match_the_header(tmp_path):
    L, _ = _lib()
    pairs = [("mobi_gemm_args", L.GemmArgs), ("mobi_attn_args", L.AttnArgs), ("mobi_groupnorm_args", L.GroupNormArgs),
             ("mobi_layernorm_args", L.LayerNormArgs), ("mobi_ln_dual_spec", L.LnDualSpec),
             ("mobi_ln_adapter_args", L.LnAdapterArgs), ("mobi_im2col_args", L.Im2colArgs),
             ("mobi_ctx_attn_args", L.CtxAttnArgs), ("mobi_sampler_args", L.SamplerArgs),
             ("mobi_assemble_args", L.AssembleArgs), ("mobi_layernorm_bwd_args", L.LayerNormBwdArgs),
             ("mobi_groupnorm_bwd_args", L.GroupNormBwdArgs), ("mobi_attn_softmax_bwd_args", L.AttnSoftmaxBwdArgs),
             ("mobi_ctx_attn_qspace_args", L.CtxAttnQspaceArgs), ("mobi_latent_input_args", L.LatentInputArgs),
             ("mobi_attn_bwd_tiles_args", L.AttnBwdTilesArgs)]
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(void){' % HEADER
    prog += "".join('printf("%%zu\\n", sizeof(%s));' % c for c, _ in pairs)
    # every field by NAME: a reordering of same-size members between the header and the binding changes no sizeof
    fields = [(c, ct, f[0]) for c, ct in pairs for f in ct._fields_]
    prog += "".join('printf("%%zu\\n", offsetof(%s, %s));' % (c, f) for c, _, f in fields) + "return 0;}\n"
    src = tmp_path / "sizes.c"
    src.write_text(prog)
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True)
    nums = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    sizes, offsets = nums[:len(pairs)], nums[len(pairs):]
    for (cname, ct), size in zip(pairs, sizes):
        assert C.sizeof(ct) == size, "%s: ctypes %d bytes, C %d bytes" % (cname, C.sizeof(ct), size)
    assert len(offsets) == len(fields) > 200
    for (cname, ct, f), off in zip(fields, offsets):
        assert getattr(ct, f).offset == off, "%s.%s: ctypes offset %d, C offset %d" % (cname, f, getattr(ct, f).offset, off)


def test_product_path_fails_loudly_without_cuda():
    """No CPU fallback: ops on CPU tensors raise instead of computing something else."""
    import torch
    from mobi_b200 import ops
    with pytest.raises(RuntimeError):
        ops.cast_bf16(torch.zeros(4))
    with pytest.raises((RuntimeError, AssertionError)):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
