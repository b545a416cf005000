"""the reference's OWN `llama_inference.load_quant` (imported unmodified from /root/reference) driven against THIS
repo's `quant` / `utils` packages -- "drops into llama_inference.py unchanged".

Only things outside the hot path are stubbed, and only in this test: `gptq` (the solver, needs texttable) and the dataset helpers the
reference imports from `utils` at module level.  Runs where the reference checkout exists (the build container); skipped elsewhere.
The forward of the loaded model is exercised on the GPU by tests/test_gpu_modules.py::test_load_quant_pipeline_on_tiny_llama.
"""
import importlib
import json
import os
import sys
import types

import pytest
import torch

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, 'llama_inference.py')), reason='reference checkout not present')


def _import_reference_llama_inference():
    import quant  # this repo's drop-in package (tests/conftest.py puts gptq-for-llama_b200 on sys.path)
    import utils as ours
    assert 'gptq-for-llama_b200' in quant.__file__ and 'gptq-for-llama_b200' in ours.__file__
    shim = types.ModuleType('utils')
    shim.find_layers, shim.DEV = ours.find_layers, ours.DEV
    for name in ('set_seed', 'get_wikitext2', 'get_ptb', 'get_c4', 'get_ptb_new', 'get_c4_new', 'get_loaders'):  # calibration data, not on the path
        setattr(shim, name, lambda *a, **k: (_ for _ in ()).throw(NotImplementedError('dataset helper outside the hot path')))
    gptq = types.ModuleType('gptq')
    gptq.GPTQ = type('GPTQ', (), {})
    saved = {k: sys.modules.get(k) for k in ('utils', 'gptq', 'llama_inference')}
    sys.modules['utils'], sys.modules['gptq'] = shim, gptq
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    try:
        sys.modules.pop('llama_inference', None)
        return importlib.import_module('llama_inference'), saved
    finally:
        sys.path.remove(REF)


def _restore(saved):
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def test_reference_load_quant_runs_against_this_quant_package(tmp_path):
    import quant
    from transformers import LlamaConfig, LlamaForCausalLM
    li, saved = _import_reference_llama_inference()
    try:
        assert li.quant is quant  # the reference module resolved `import quant` to this repo's package
        cfg = LlamaConfig(hidden_size=256, intermediate_size=768, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=2, vocab_size=320,
                          max_position_embeddings=64, rms_norm_eps=1e-6)
        cfg.save_pretrained(tmp_path)
        # a GPTQ checkpoint with the reference's state_dict keys / shapes / dtypes: built by the same make_quant_linear recipe, random packed tensors
        torch.set_default_dtype(torch.half)
        src = LlamaForCausalLM(cfg)
        torch.set_default_dtype(torch.float)
        layers = li.find_layers(src)
        layers.pop('lm_head', None)
        quant.make_quant_linear(src, layers, 4, 128)
        g = torch.Generator().manual_seed(0)
        for m in src.modules():
            if isinstance(m, quant.QuantLinear):
                m.qweight.copy_(torch.randint(-2**31, 2**31 - 1, m.qweight.shape, generator=g, dtype=torch.int64).to(torch.int32))
                m.qzeros.copy_(torch.randint(-2**31, 2**31 - 1, m.qzeros.shape, generator=g, dtype=torch.int64).to(torch.int32))
                m.scales.copy_((torch.rand(m.scales.shape, generator=g) * 1e-2 + 1e-3).half())
        sd = src.state_dict()
        assert sd['model.layers.0.self_attn.q_proj.qweight'].shape == (256 // 8, 256) and sd['model.layers.0.mlp.down_proj.qzeros'].dtype == torch.int32
        ckpt = os.path.join(tmp_path, 'tiny-4bit-128g.pt')
        torch.save(sd, ckpt)

        model = li.load_quant(str(tmp_path), ckpt, 4, 128)  # the reference's function, unmodified: fused attention, Triton-named norm, fused MLP, warm-up
        assert model.seqlen == 2048 and not model.training
        layer = model.model.layers[0]
        assert isinstance(layer.self_attn, quant.QuantLlamaAttention) and isinstance(layer.mlp, quant.QuantLlamaMLP)
        assert isinstance(layer.input_layernorm, quant.TritonLlamaRMSNorm) and isinstance(model.model.norm, quant.TritonLlamaRMSNorm)
        # the checkpoint tensors arrived where the kernels read them: fused q|k|v along N, gate/up buffers of the fused MLP, down_proj untouched
        q, k, v = (sd[f'model.layers.0.self_attn.{n}_proj.qweight'] for n in 'qkv')
        assert torch.equal(layer.self_attn.qkv_proj.qweight, torch.cat([q, k, v], dim=1))
        assert torch.equal(layer.mlp.gate_proj_qweight, sd['model.layers.0.mlp.gate_proj.qweight'])
        assert torch.equal(layer.mlp.up_proj_scales, sd['model.layers.0.mlp.up_proj.scales'])
        assert torch.equal(layer.mlp.down_proj.qzeros, sd['model.layers.0.mlp.down_proj.qzeros'])
        assert torch.equal(model.lm_head.weight, sd['lm_head.weight'])  # never quantized (llama_inference.py:46-48)
        with open(os.path.join(tmp_path, 'config.json')) as f:
            assert json.load(f)['hidden_size'] == 256
    finally:
        torch.set_default_dtype(torch.float)
        _restore(saved)
