"""The reference-facing API mirror on a real GPU: init_model / decode_one_token_ar / decode_n_tokens / generate /
generate_streaming (fish_tts/models/inference.py) and FishTTS / VoiceProfile / get_instance (fish_tts/synthesizer.py)."""
import numpy as np
import pytest
import torch

from fish_tts_b200.config import tiny_config
from fish_tts_b200.synthetic import make_state_dict, random_voice_codes, synthetic_prompt
from oracle import ref_harness as rh

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model_dir(tmp_path_factory):
    cfg = tiny_config(max_seq_len=4096)      # generate_long refuses prompts longer than max_seq_len - 2048 (inference.py:794)
    d = tmp_path_factory.mktemp("model")
    rh.fabricate_model_dir(cfg, make_state_dict(cfg, seed=0), d)
    return cfg, d


def test_inference_mirror_end_to_end(model_dir):
    from fish_tts_b200 import inference as inf
    cfg, d = model_dir
    model, decode_one_token = inf.init_model(str(d), "cuda", torch.bfloat16, compile=True)
    assert model.config.semantic_begin_id == cfg.semantic_begin_id and model.tokenizer.get_token_id("<|im_end|>") == cfg.im_end_id
    prompt = synthetic_prompt(cfg, 5, 12, 4, seed=1).cuda()
    T, C1 = prompt.size(1), cfg.num_codebooks + 1
    kw = dict(temperature=0.7, top_p=0.8, repetition_penalty=1.1)
    # (1) engine-owned loop
    seq = inf.generate(model=model, prompt=prompt, max_new_tokens=20, audio_masks=None, audio_parts=None, decode_one_token=decode_one_token, **kw)
    assert seq.shape == (C1, T + 20) and torch.equal(seq[:, :T], prompt)
    # (2) streaming yields the same code columns (semantic row dropped)
    cols = list(inf.generate_streaming(model=model, prompt=prompt, max_new_tokens=20, audio_masks=None, audio_parts=None,
                                       decode_one_token=decode_one_token, **kw))
    assert len(cols) == 20 and all(c.shape == (cfg.num_codebooks, 1) for c in cols)
    assert torch.equal(torch.cat(cols, dim=1), seq[1:, T:])
    # (3) the reference-style host loop (prefill call + decode_n_tokens + per-step callable) gives the same columns bit for bit:
    #     same engine, same Philox stream; this checks the device-side window / position bookkeeping against the host version
    t = [torch.tensor(v, device="cuda", dtype=torch.float) for v in (0.7, 0.8, 1.1)]
    first = inf.decode_one_token_ar(model, prompt.view(1, C1, -1), torch.arange(T, device="cuda"), *t, None, None).clone()
    rest = inf.decode_n_tokens(model, first.view(1, C1, -1), torch.tensor([T], device="cuda", dtype=torch.int), 19, *t, None, None,
                               decode_one_token=decode_one_token)
    host = torch.cat([first, rest], dim=1)
    assert torch.equal(host, seq[:, T:])
    with pytest.raises(ValueError, match="exceeds max_seq_len"):
        inf.generate(model=model, prompt=torch.zeros((C1, cfg.max_seq_len), dtype=torch.int32, device="cuda"), max_new_tokens=4)
    model.engine.close()


def test_fishtts_api(model_dir):
    from fish_tts_b200 import synthesizer as syn
    cfg, d = model_dir
    with pytest.raises(RuntimeError):
        syn.FishTTS(d, device="cpu")

    def prompt_encoder(texts, codes, text):      # stands in for ContentSequence.encode_for_inference (out of scope)
        n = sum(c.shape[1] for c in codes)
        p = synthetic_prompt(cfg, 3, 0, 3, seed=len(text)).numpy()
        if n:
            vq = np.concatenate([np.asarray(c) for c in codes], axis=1)
            mid = np.zeros((cfg.num_codebooks + 1, n), dtype=np.int32)
            mid[0], mid[1:] = vq[0] + cfg.semantic_begin_id, vq
            p = np.concatenate([p[:, :3], mid, p[:, 3:]], axis=1)
        return p

    def vocoder(codes):                           # 2048 samples per frame, like the DAC (vocoder.py:854, 872)
        return np.zeros(int(codes.shape[1]) * 2048, dtype=np.float32)

    syn.reset_instance()
    tts = syn.get_instance(model_dir=d, prompt_encoder=prompt_encoder, vocoder=vocoder)
    assert syn.get_instance() is tts and tts.sample_rate == 44100 and tts.precision == "bf16"
    prof = syn.VoiceProfile(codes=random_voice_codes(cfg, 30).numpy(), text="ref", name="a")
    tts.set_references([prof]); tts.add_reference(prof)
    assert tts.num_references == 2 and len(tts.get_references()) == 2
    tts.clear_references(); assert tts.num_references == 0
    tts.set_references([prof])
    codes = tts.generate_codes("hello", max_tokens=24)
    assert codes.shape == (cfg.num_codebooks, 23)           # the caller drops the last column (inference.py:839)
    wav = tts.synthesize("hello", max_tokens=24)
    assert wav[:4] == b"RIFF"
    chunks = list(tts.synthesize_stream("hello", max_tokens=45, chunk_tokens=20, min_first_chunk=10))
    assert sum(len(c) for c in chunks) == 45 * 2048 * 2 and len(chunks) == 3     # 10 + 20 + 15 frames
    tts._vocoder = None
    with pytest.raises(RuntimeError, match="Vocoder not loaded"):
        tts.synthesize("x")
    tts._model.engine.close()
    syn.reset_instance()


def test_streaming_hand_off_and_prefix_reuse(model_dir):
    """DualAREngine.stream (dualar_decode_async: chunked copies into pinned buffers behind events, the next chunk enqueued before
    the host waits) yields exactly the columns generate() produces; and a second request whose prompt shares the VoiceProfile
    prefix re-prefills only the differing tail (prefix_reuse) with bit-identical tokens."""
    from fish_tts_b200.engine import DualAREngine
    cfg, d = model_dir
    sd = make_state_dict(cfg, seed=0)
    eng = DualAREngine(cfg, sd, device=0, seed=9)
    S = dict(temperature=0.7, top_p=0.8, repetition_penalty=1.1)
    base = synthetic_prompt(cfg, 5, 120, 0, seed=1)                      # the "prefilled VoiceProfile" part
    tails = [synthetic_prompt(cfg, 6 + i, 0, 0, seed=20 + i) for i in range(3)]
    prompts = [torch.cat([base, t], dim=1) for t in tails]
    # (1) streaming == one-shot
    want = eng.generate(prompts[0], 57, **S)
    chunks = list(eng.stream(prompts[0], 57, **S, first_chunk=10, chunk=20))
    assert [c.shape[1] for c in chunks] == [10, 20, 20, 7]
    assert (np.concatenate(chunks, axis=1) == want).all()
    assert int(eng.read("prefix_reused")[0]) == prompts[0].size(1) - 1    # the same prompt again: everything but the last position is reused
    # (2) prefix reuse across different utterances of the same voice
    outs = {}
    for reuse in (1, 0):
        eng.set_option("prefix_reuse", reuse)
        eng.set_option("prefill_mode", 0)                                  # (marks the cache dirty: the first request prefills everything)
        res, kept = [], []
        for p in prompts:
            res.append(eng.generate(p, 16, **S))
            kept.append(int(eng.read("prefix_reused")[0]))
        outs[reuse] = (res, kept)
    assert outs[0][1] == [0, 0, 0] and outs[1][1][0] == 0 and min(outs[1][1][1:]) >= base.size(1)
    for a, b in zip(outs[0][0], outs[1][0]):
        assert (a == b).all(), "tokens after a prefix-reusing prefill differ from a full prefill"
    # (3) EOS inside a chunk ends the stream with exactly the columns up to <|im_end|>
    eng.close()
    eng = DualAREngine(cfg, make_state_dict(cfg, seed=0, eos_reachable=True), device=0, seed=3)
    for trial in range(4):
        eng.seed(50 + trial)
        want = eng.generate(prompts[1], 64, **S)
        eng.seed(50 + trial)
        got = np.concatenate(list(eng.stream(prompts[1], 64, **S)), axis=1)
        assert got.shape == want.shape and (got == want).all()
    eng.close()


def test_install_beneath_the_reference_fishtts(model_dir):
    """fish_tts_b200.inference.install() followed by the REFERENCE's own get_instance / set_references /
    synthesize / synthesize_stream (synthesizer.py:161-167, 363-377, 431-584), with a stub vocoder (codec.pth is out of scope and
    `dac` / `audiotools` are not installed).  The reference class runs unchanged on the engine; its own weights are released."""
    if not rh.reference_available():
        pytest.skip("no copy of the reference (oracle/make_ref.py) on this box")
    cfg, d = model_dir
    _, ref_inf = rh.import_reference()
    import fish_tts.synthesizer as rsyn
    from fish_tts_b200 import inference as inf
    inf.install(ref_inf)
    try:
        rsyn.reset_instance()
        tts = rsyn.get_instance(model_dir=str(d), device="cuda", precision="bf16", warmup=True)
        assert tts is rsyn.get_instance() and hasattr(tts._model, "_dualar_engine")
        assert sum(p.numel() for p in tts._model.parameters()) == 0, "the reference module's weights must be released after the engine is built"

        class StubVocoder:      # DAC.decode(indices, feature_lengths) -> (audio (1, 1, n * 2048), lengths)   vocoder.py:906-912
            def decode(self, codes, feature_lengths):
                assert codes.dim() == 3 and codes.size(1) == cfg.num_codebooks and int(codes.min()) >= 0
                return torch.zeros((1, 1, codes.size(-1) * 2048), device=codes.device), feature_lengths

        tts._vocoder = StubVocoder()
        prof = rsyn.VoiceProfile(codes=random_voice_codes(cfg, 40).numpy(), text="reference transcript", name="a")
        tts.set_references([prof])
        assert tts.num_references == 1
        wav = tts.synthesize("hello world", max_tokens=30)
        assert wav[:4] == b"RIFF" and len(wav) == 44 + 29 * 2048 * 2        # the caller drops the last column (inference.py:839)
        chunks = list(tts.synthesize_stream("hello again", max_tokens=45, chunk_tokens=20, min_first_chunk=10))
        assert sum(len(c) for c in chunks) == 45 * 2048 * 2
        kept = int(tts._model._dualar_engine.read("prefix_reused")[0])
        assert kept >= 40, f"the second utterance of the same voice must reuse the VoiceProfile's KV rows (kept {kept})"
        tts._model._dualar_engine.close()
    finally:
        inf.uninstall(ref_inf)
        rsyn.reset_instance()
