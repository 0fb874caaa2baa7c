"""CPU: the host front end (text -> ids, duration heuristic, wav I/O) against the reference's own functions, AST-extracted from
F5_TTS/F5-TTS-ONNX-Inference.py where it lies (skipped where /root/reference is absent: the GPU box), and against fixed cases."""
import ast
import os
import re
import types

import numpy as np
import pytest

import b200tts  # noqa: F401
from b200tts import frontend as fe

REF = "/root/reference/F5_TTS/F5-TTS-ONNX-Inference.py"

VOCAB = [" ", "a", "b", "c", "d", "e", "h", "l", "o", "r", "w", ",", ".", "'", "ni3", "hao3", "shi4", "jie4", "!", "x"]


def _write_vocab(tmp_path):
    p = tmp_path / "vocab.txt"
    p.write_text("".join(s + "\n" for s in VOCAB), encoding="utf-8")
    return str(p)


def _toy_pinyin(s):
    table = {"你": "ni3", "好": "hao3", "世": "shi4", "界": "jie4"}
    return [table.get(ch, ch) for ch in s]


def _reference_functions():
    """convert_char_to_pinyin / list_str_to_idx compiled from the reference script's own source (the script has top-level side
    effects, so it cannot be imported), with jieba / pypinyin replaced by the same toy segmenter / pinyin table the test feeds
    to the port."""
    import torch
    src = open(REF, encoding="utf-8").read()
    tree = ast.parse(src)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("convert_char_to_pinyin", "list_str_to_idx")]
    assert len(wanted) == 2
    jieba = types.SimpleNamespace(dt=types.SimpleNamespace(initialized=True), cut=fe._fallback_segments)
    ns = {"jieba": jieba, "lazy_pinyin": lambda s, style=None, tone_sandhi=True: _toy_pinyin(s), "Style": types.SimpleNamespace(TONE3=3),
          "torch": torch}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), REF, "exec"), ns)
    return ns["convert_char_to_pinyin"], ns["list_str_to_idx"]


TEXTS = ["hello world", "hello,world. it's", "你好,world", "abc你好世界!", "a;b “c” ‘d’", "  lead", "x你y"]


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")
def test_text_to_ids_equals_the_reference_functions(tmp_path):
    ref_convert, ref_to_idx = _reference_functions()
    vocab = fe.load_vocab(_write_vocab(tmp_path))
    for t in TEXTS:
        want = ref_convert([t])
        got = fe.text_to_symbols([t], segmenter=fe._fallback_segments, to_pinyin=_toy_pinyin)
        assert got == want, (t, got, want)
    want_ids = ref_to_idx(ref_convert(TEXTS), vocab).numpy()
    got_ids = fe.symbols_to_ids(fe.text_to_symbols(TEXTS, segmenter=fe._fallback_segments, to_pinyin=_toy_pinyin), vocab)
    assert got_ids.dtype == np.int32
    np.testing.assert_array_equal(got_ids, want_ids)


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")
def test_text_to_symbols_random_strings_equal_the_reference(tmp_path):
    """2 000 random strings over ASCII words, digits, quotes, CJK with and without toy pinyin, full-width and other non-ASCII
    characters, through the port and through the reference's own function (both with the same segmenter / pinyin stand-ins)."""
    import random
    ref_convert, ref_to_idx = _reference_functions()
    vocab = fe.load_vocab(_write_vocab(tmp_path))
    rng = random.Random(17)
    alphabet = ["a", "b", "hello", "World", " ", "  ", ",", ".", ";", ":", "'", '"', "“", "”", "‘", "’", "1", "23", "-", "!", "你", "好", "世", "界", "吗",
                "，", "。", "é", "ü", "Ω", "ｱ", "？", "\n", "\t", "it's", "x"]
    for polyphone in (True, False):
        for _ in range(1000):
            t = "".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 14)))
            want = ref_convert([t], polyphone=polyphone)
            got = fe.text_to_symbols([t], polyphone=polyphone, segmenter=fe._fallback_segments, to_pinyin=_toy_pinyin)
            assert got == want, (t, polyphone, got, want)
            np.testing.assert_array_equal(fe.symbols_to_ids(got, vocab), ref_to_idx(want, vocab).numpy())


def test_text_to_symbols_fixed_cases(tmp_path):
    vocab = fe.load_vocab(_write_vocab(tmp_path))
    assert vocab[" "] == 0 and vocab["x"] == len(VOCAB) - 1
    syms = fe.text_to_symbols(["hello,world"], segmenter=fe._fallback_segments, to_pinyin=_toy_pinyin)[0]
    assert "".join(syms) == "hello, world"                      # a word glued to punctuation gets a separating space
    syms = fe.text_to_symbols(["你好"], segmenter=lambda t: [t], to_pinyin=_toy_pinyin)[0]
    assert syms == [" ", "ni3", " ", "hao3"]
    ids = fe.symbols_to_ids([["a", "?", "b"], ["c"]], vocab)
    np.testing.assert_array_equal(ids, [[1, 0, 2], [3, -1, -1]])  # unknown -> 0, padding -1


def test_duration_heuristic_and_the_pause_regex_quirk():
    # reference arithmetic, F5-TTS-ONNX-Inference.py:227-231
    ref, gen, L = "hello world.", "this is a test of it", 144000
    frames = L // 256 + 1
    assert fe.estimate_max_duration(ref, gen, L) == frames + int(frames / len(ref) * len(gen) / 1.0)
    assert fe.estimate_max_duration(ref, gen, L, speed=2.0) == frames + int(frames / len(ref) * len(gen) / 2.0)
    # q13: single pause marks do NOT add 3 -- only the literal seven-character run does
    zh = "你好，世界。"
    assert len(re.findall(fe._ZH_PAUSE_PATTERN, zh)) == 0
    assert fe.estimate_max_duration(zh, zh, L) == frames + frames
    odd = "a。，、；：？！b"
    assert fe.estimate_max_duration(odd, "a", L) == frames + int(frames / (len(odd.encode()) + 3) * 1)


def test_wav_round_trip_and_channel_mix(tmp_path):
    rng = np.random.default_rng(0)
    pcm = rng.integers(-20000, 20000, size=5000, dtype=np.int16)
    p = str(tmp_path / "a.wav")
    fe.save_wav(p, pcm, 24000)
    np.testing.assert_array_equal(fe.load_wav_mono_int16(p), pcm)
    import wave
    st = np.stack([pcm, pcm[::-1]], 1)
    p2 = str(tmp_path / "st.wav")
    with wave.open(p2, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(24000); w.writeframes(st.astype("<i2").tobytes())
    mono = fe.load_wav_mono_int16(p2)
    np.testing.assert_array_equal(mono, np.floor((pcm.astype(np.float64) + pcm[::-1]) / 2).astype(np.int16))
    p3 = str(tmp_path / "r.wav")
    fe.save_wav(p3, pcm, 48000)
    half = fe.load_wav_mono_int16(p3)                             # 48 kHz -> 24 kHz
    assert half.size == 2500 and np.abs(half.astype(np.int32) - pcm[::2]).max() <= 1
