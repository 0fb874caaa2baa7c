"""GPU parity for the IndexTTS sentence loop (Inference_IndexTTS_ONNX.py:719-804): text -> sentences -> per
sentence the full-size GPT-2 greedy decode and the IndexTTS_F vocoder through `IndexTTSSynthesizer`, against the oracle driven by
the same host loop restated here (penalty vector carried from sentence to sentence, 200 ms of silence after every sentence).

Stated tolerances, fp32 engine: mel ids identical (the oracle's smallest top-1/top-2 logit gap over these calls is asserted to be
>= 1e-2, fp32 round-off through 24 layers is ~1e-3), PCM SNR >= 40 dB against the oracle's waveform (hidden rows agree to 5e-3 on
magnitude ~5), silence bit-exact."""
import numpy as np
import pytest

import b200tts  # noqa: F401
from b200tts import capi, config, synth, weights
from b200tts import indextts_frontend as fe
from conftest import snr_db
from oracle import indextts_gpt_ref as RG
from oracle import indextts_ref as RV

GPT, VOC = config.INDEXTTS_GPT, config.INDEXTTS_VOCODER
TEXT = "大家好。hello world! 你好世界?"
MAX_NEW = 6
SEED_GPT, SEED_VOC, SEED_IN = 556, 777, 78


@pytest.fixture(scope="module")
def tokenizer(tmp_path_factory):
    import sentencepiece as spm
    d = tmp_path_factory.mktemp("bpe")
    lines = ["HELLO WORLD , THIS IS A TEST .", "你 好 世 界 , 大 家 好 .", "I AM HERE ! ARE YOU THERE ?"]
    (d / "corpus.txt").write_text("\n".join(lines * 40), encoding="utf-8")
    spm.SentencePieceTrainer.train(input=str(d / "corpus.txt"), model_prefix=str(d / "bpe"), vocab_size=70, model_type="bpe",
                                   character_coverage=1.0, bos_id=0, eos_id=1, unk_id=2, minloglevel=2)
    ident = fe.IdentityNormalizer()
    return fe.TextTokenizer(str(d / "bpe.model"), fe.TextNormalizer(ident, ident))


def oracle_sentence_loop(tokenizer, conditioning, text, max_tokens_per_sentence):
    """-> (waveforms per sentence incl. the pad, mel ids per sentence, smallest top-2 logit gap)."""
    sd_g, sd_v = synth.igpt_state(SEED_GPT, GPT), synth.ivgan_state(SEED_VOC)
    conds, cond_layer, latent = conditioning[:-2], conditioning[-2], conditioning[-1]
    pen = np.ones((1, GPT.mel_codes), np.float32)
    pad = np.zeros((1, 1, int(24000 * 0.2)), np.int16)
    waves, mel_ids, gap = [], [], np.inf
    for piece in tokenizer.split_sentences(tokenizer.tokenize(text), max_tokens_per_sentence):
        ids = np.asarray([tokenizer.convert_tokens_to_ids(piece)], np.int32)
        got, hidden, pen, logits = RG.generate(latent, ids, sd_g, GPT, max_new=MAX_NEW, penalty=pen, return_logits=True)
        top = np.sort(logits, axis=1)
        gap = min(gap, float((top[:, -1] - top[:, -2]).min()))
        pcm, wavef = RV.indextts_f_pcm(hidden, conds, cond_layer, sd_v, VOC, return_float=True)
        waves.append((np.concatenate([pcm.numpy().reshape(1, 1, -1), pad], axis=-1), wavef.numpy().reshape(-1)))
        mel_ids.append(got)
    return waves, mel_ids, gap


def make_conditioning():
    conds, cond_layer, _ = synth.ivgan_inputs(301, 3)
    latent, _ = synth.igpt_inputs(SEED_IN, 4, GPT)
    return list(conds) + [cond_layer, latent]


def test_oracle_side_is_decisive(tokenizer):
    """CPU: the oracle half of the GPU test below -- three sentences, a top-2 gap that fp32 round-off cannot flip."""
    waves, mel_ids, gap = oracle_sentence_loop(tokenizer, make_conditioning(), TEXT, 6)
    assert len(waves) == 3 and all(len(m) == MAX_NEW for m in mel_ids)
    assert gap >= 1e-2, gap
    assert all(w.shape[-1] == 1024 * (MAX_NEW - 2) + 30 + 4800 for w, _ in waves)


@pytest.mark.gpu
def test_sentence_loop_vs_oracle(engine, tokenizer):
    engine.load_state("igpt", weights.igpt_engine_tensors(synth.igpt_state(SEED_GPT, GPT), GPT))
    engine.indextts_gpt_build()
    engine.load_state("ivgan", weights.ivgan_engine_tensors(synth.ivgan_state(SEED_VOC), VOC))
    engine.indextts_vocoder_build()
    conditioning = make_conditioning()
    waves, mel_ids, gap = oracle_sentence_loop(tokenizer, conditioning, TEXT, 6)
    assert gap >= 1e-2
    syn = fe.IndexTTSSynthesizer(engine, tokenizer, precision=capi.F32, max_tokens_per_sentence=6)
    whole = syn.synthesize(conditioning, TEXT, keep="all", max_new=MAX_NEW)
    assert whole.dtype == np.int16 and whole.shape == (1, 1, sum(w.shape[-1] for w, _ in waves))
    assert [t["mel_ids"].tolist() for t in syn.trace] == [m.tolist() for m in mel_ids]
    at = 0
    for (want, wantf), t in zip(waves, syn.trace):
        n = t["samples"]
        got = whole[0, 0, at:at + n]
        assert snr_db(want[0, 0, :n], got) >= 40.0
        assert not whole[0, 0, at + n:at + want.shape[-1]].any()                  # the 200 ms split pad
        at += want.shape[-1]
    last = syn.synthesize(conditioning, TEXT, max_new=MAX_NEW)                     # keep="last": the file the reference script writes
    np.testing.assert_array_equal(last, whole[..., -waves[-1][0].shape[-1]:])      # same penalty carry -> the same last sentence
    bf = fe.IndexTTSSynthesizer(engine, tokenizer, max_tokens_per_sentence=6).synthesize(conditioning, TEXT, keep="all", max_new=MAX_NEW)
    assert bf.shape == whole.shape                                                 # the default (bf16) engine runs the same loop
