"""Compatibility shim that lets the UNMODIFIED reference (qwopqwop200/GPTQ-for-LLaMa, triton branch @ e985b70) import and run
under the Triton 3.x / torch 2.11 of this image.  The reference was written against Triton 2.0:

  * quant/custom_autotune.py:72   triton.testing.do_bench(..., percentiles=(0.5, 0.2, 0.8), rep=40)   -> `quantiles=`
  * quant/custom_autotune.py:73   triton.compiler.OutOfResources                                      -> triton.runtime.errors
  * quant/fused_attn.py:43        tl.libdevice.exp                                                    -> triton.language.extra.libdevice

Nothing in the reference is edited: import this module BEFORE `import quant` (the reference tree, or its copy under the
git-ignored baseline/_ref/).  It is benchmark / test tooling, never imported by the product package.
"""
import triton
import triton.language as tl
import triton.testing

if not hasattr(triton.compiler, 'OutOfResources'):
    from triton.runtime.errors import OutOfResources
    triton.compiler.OutOfResources = OutOfResources

if not hasattr(tl, 'libdevice'):
    from triton.language.extra import libdevice as _libdevice
    tl.libdevice = _libdevice

_do_bench = triton.testing.do_bench


def _do_bench_compat(fn, warmup=25, rep=100, grad_to_none=None, percentiles=None, quantiles=None, **kw):
    q = quantiles if quantiles is not None else percentiles
    kw.pop('fast_flush', None)
    return _do_bench(fn, warmup=warmup, rep=rep, grad_to_none=grad_to_none, quantiles=q, **kw)


triton.testing.do_bench = _do_bench_compat
