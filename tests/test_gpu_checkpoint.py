"""GPU side of the checkpoint path: the offline quantiser writes files byte-identical to the reference's
quantisers (Q4ByteBufferTensor.java:66-120, Q8ByteBufferTensor.java:47-90, SafeTensorSupport.quantizeModel :215-332), and a
checkpoint directory loaded through the C ABI generates exactly what the same weights handed over in memory generate."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _f32_checkpoint(tmp_path, cfg, dtype):
    from jlama_b200 import native, synth
    from jlama_b200 import safetensors_io as sio
    w = synth.make_weights(cfg, wdtype=dtype, embed_dtype=dtype)
    src = tmp_path / "src"
    sio.save_checkpoint(str(src), w, cfg)
    return w, src


@pytest.mark.parametrize("src_dtype", ["F32", "BF16"])
@pytest.mark.parametrize("qname", ["Q4", "I8"])
def test_quantize_model_is_byte_identical_to_the_reference_quantisers(cuda_ctx, oracle, tmp_path, src_dtype, qname):
    from jlama_b200 import native, synth
    from jlama_b200 import safetensors_io as sio
    cfg = synth.get_config("tiny")
    sd = native.F32 if src_dtype == "F32" else native.BF16
    qt = native.Q4 if qname == "Q4" else native.I8
    w, src = _f32_checkpoint(tmp_path, cfg, sd)
    dst = sio.quantize_model(cuda_ctx, str(src), str(tmp_path / "dst"), qt)
    with sio.SafeTensors(dst) as st:
        assert st.majority_dtype() == qt
        for name, (dt, data, _) in w.items():
            f = oracle.bf16_to_f32(data) if dt == native.BF16 else data
            if "norm" in name:  # default skip pattern (QuantizeCommand.java:36-38) and 1-row tensors keep their dtype
                assert st.info(name)["dtype_code"] == dt and np.array_equal(st.get(name), data)
                continue
            dt2, q2, s2 = st.load(name)
            assert dt2 == qt, name
            q_ref, s_ref = (oracle.quantize_q4 if qt == native.Q4 else oracle.quantize_q8_weights)(f)
            assert np.array_equal(q2, q_ref), name
            assert np.array_equal(s2.view(np.uint32), s_ref.view(np.uint32)), name
    assert (tmp_path / "dst" / "config.json").exists()


def test_checkpoint_directory_loads_and_generates_like_in_memory_weights(cuda_ctx, oracle, tmp_path):
    from jlama_b200 import synth
    from jlama_b200 import safetensors_io as sio
    from jlama_b200.model import LlamaModel
    cfg = synth.get_config("small")
    w = synth.make_weights(cfg)
    sio.save_checkpoint(str(tmp_path / "ckpt"), w, cfg)
    a = LlamaModel.from_checkpoint(cuda_ctx, str(tmp_path / "ckpt"))
    b = LlamaModel(cuda_ctx, cfg, w)
    prompt = synth.random_prompt(cfg, 13)
    ta, la = a.generate(prompt, 12, want_logits=True)
    tb, lb = b.generate(prompt, 12, want_logits=True)
    assert list(ta) == list(tb) and np.array_equal(la, lb)
    ot, _ = oracle.OracleLlama(cfg, w, act_q8=True).generate(prompt, 12)
    assert list(ta) == list(ot)
    a.close()
    b.close()


def test_mixtral_checkpoint_directory_loads_like_in_memory_weights(cuda_ctx, oracle, tmp_path):
    """MixtralModel.java:88-105 tensor names (block_sparse_moe.gate / experts.<e>.w1|w2|w3) and MixtralConfig's num_local_experts /
    num_experts_per_tok read behind the C ABI: same registrations as the in-memory path, so the generations are bit-identical."""
    from jlama_b200 import synth
    from jlama_b200 import safetensors_io as sio
    from jlama_b200.model import LlamaModel
    cfg = synth.get_config("tiny-mixtral")
    w = synth.make_weights(cfg)
    sio.save_checkpoint(str(tmp_path / "mx"), w, cfg)
    a = LlamaModel.from_checkpoint(cuda_ctx, str(tmp_path / "mx"))
    assert a.cfg["experts"] == 8 and a.cfg["experts_per_token"] == 2
    b = LlamaModel(cuda_ctx, cfg, w)
    prompt = synth.random_prompt(cfg, 9)
    ta, la = a.generate(prompt, 8, want_logits=True)
    tb, lb = b.generate(prompt, 8, want_logits=True)
    assert list(ta) == list(tb) and np.array_equal(la, lb)
    a.close()
    b.close()
