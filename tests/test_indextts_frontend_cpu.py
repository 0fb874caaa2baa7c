"""CPU: the IndexTTS host front end (CJK pre-tokenisation, text normaliser, SentencePiece tokenizer, sentence splitter, sentence
loop) against the reference's own classes, AST-extracted from IndexTTS/Inference_IndexTTS_ONNX.py where it lies (skipped where
/root/reference is absent: the GPU box), and against fixed cases that hold everywhere."""
import ast
import os
import random
import sys
import warnings

import numpy as np
import pytest

import b200tts  # noqa: F401
from b200tts import indextts_frontend as fe

REF = "/root/reference/IndexTTS/Inference_IndexTTS_ONNX.py"
needs_ref = pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")


class ToyVerbaliser:
    """Stand-in for WeTextProcessing (not installable offline), fed to BOTH sides: digits are spoken, so an unprotected
    pinyin tone digit or a changed call order would show."""

    def __init__(self, words):
        self.words = words

    def normalize(self, text):
        return "".join(self.words.get(ch, ch) for ch in text)


ZH = ToyVerbaliser({"1": "一", "2": "二", "3": "三", "4": "四", "5": "五", "%": "百分之"})
EN = ToyVerbaliser({"1": " one ", "2": " two ", "3": " three ", "4": " four ", "5": " five "})


def _reference_namespace():
    """The four definitions compiled from the reference script's own source (the script builds ORT sessions at import time, so it
    cannot be imported)."""
    import functools
    import platform
    import re
    import traceback
    import typing
    from sentencepiece import SentencePieceProcessor
    tree = ast.parse(open(REF, encoding="utf-8").read())
    names = ("tokenize_by_CJK_char", "de_tokenized_by_CJK_char", "TextNormalizer", "TextTokenizer")
    wanted = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    assert len(wanted) == 4
    ns = {"re": re, "os": os, "warnings": warnings, "platform": platform, "traceback": traceback, "lru_cache": functools.lru_cache,
          "List": typing.List, "Union": typing.Union, "overload": typing.overload, "SentencePieceProcessor": SentencePieceProcessor}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), REF, "exec"), ns)
    return ns


def _reference_normalizer(ns):
    norm = ns["TextNormalizer"]()
    norm.zh_normalizer, norm.en_normalizer = ZH, EN
    norm.load = lambda: None                             # `load` imports tn / wetext
    return norm


@pytest.fixture(scope="module")
def bpe_model(tmp_path_factory):
    """A small BPE model trained here (the real bpe.model ships with the IndexTTS checkpoint, which is not on disk)."""
    import sentencepiece as spm
    d = tmp_path_factory.mktemp("bpe")
    lines = ["HELLO WORLD , THIS IS A TEST .", "你 好 世 界 , 大 家 好 .", "I AM HERE ! ARE YOU THERE ?", "现 在 正 在 体 验 AI 科 技 .",
             "XVAN2 ZE2 , ONE - TWO - THREE … ' QUOTED ' .", "恩 , 母 亲 说 : 一 二 三 四 五 ."]
    corpus = d / "corpus.txt"
    corpus.write_text("\n".join(lines * 40), encoding="utf-8")
    spm.SentencePieceTrainer.train(input=str(corpus), model_prefix=str(d / "bpe"), vocab_size=120, model_type="bpe",
                                   character_coverage=1.0, bos_id=0, eos_id=1, unk_id=2, minloglevel=2)
    return str(d / "bpe.model")


CJK_LINES = ["你好世界是 hello world 的中文", "  ", "", "abc", "한국어 text ｶﾅ mixed 𠀀 end", "a  b\tc你", "See you!  好的", "ＡＢＣ全角"]
DETOK_LINES = ["你 好 世 界 是 HELLO WORLD 的 中 文", "SEE YOU!", "你 好", "A,B 好", "HELLO-WORLD 好 FOO BAR", "x sent s 好 e", "", "a b  c",
               "MR. SMITH 来 了 , OK"]


def test_cjk_tokenisation_fixed_cases():
    assert fe.tokenize_by_CJK_char("你好世界是 hello world 的中文") == "你 好 世 界 是 HELLO WORLD 的 中 文"
    assert fe.tokenize_by_CJK_char("a 你b", do_upper_case=False) == "a 你 b"
    assert fe.tokenize_by_CJK_char("   ") == ""
    assert fe.de_tokenized_by_CJK_char("你 好 世 界 是 HELLO WORLD 的 中 文") == "你好世界是HELLO WORLD的中文"
    assert fe.de_tokenized_by_CJK_char("SEE YOU!", do_lower_case=True) == "see you!"
    assert fe.de_tokenized_by_CJK_char("A,B") == "A,<sent_1>"                  # quirk i6


@needs_ref
def test_cjk_tokenisation_equals_the_reference():
    ns = _reference_namespace()
    for line in CJK_LINES + DETOK_LINES:
        for flag in (True, False):
            assert fe.tokenize_by_CJK_char(line, flag) == ns["tokenize_by_CJK_char"](line, flag), line
            assert fe.de_tokenized_by_CJK_char(line, flag) == ns["de_tokenized_by_CJK_char"](line, flag), line
    rng = random.Random(5)
    alphabet = "ab AB-你好ｶ한,.!<>_0sent"
    for _ in range(400):
        line = "".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 24)))
        assert fe.tokenize_by_CJK_char(line) == ns["tokenize_by_CJK_char"](line), line
        assert fe.de_tokenized_by_CJK_char(line) == ns["de_tokenized_by_CJK_char"](line), line


NORM_TEXTS = [
    "大家好，我现在正在大可奇奇体验 ai 科技。", "Hello: world; it costs 3 (three) dollars... ok~", "晕 xuan4 是一种 gan3 觉，qu4 ba", "克里斯托弗·诺兰 和 约翰-保罗 来了",
    "test@example.com", "12345", "嗯，呣呣！", "“curly” ‘quotes’ \"straight\" 【括号】《书名》", "价格是$5，，，然后……完", "a, \"'\"), (b", "line one\nline two ,,, done",
    "ju2 zi5 lüe4 xun1 JUAN3", "I have 2 cats: Tom — and Jerry [sic]", "   trailing spaces   ", "",
]


def test_normaliser_fixed_cases():
    norm = fe.TextNormalizer(ZH, EN)
    norm.load()                                           # keeps the injected verbalisers
    assert norm.normalize("晕 xuan4 是一种 gan3 觉") == "晕 XVAN4 是一种 gan3 觉"     # tone digits protected, j/q/x + u -> V
    assert norm.normalize("有2个") == "有二个"
    assert norm.normalize("Hello: 2 (two)") == "Hello,  two  'two'"
    assert norm.normalize("“x”") == "“x”"                                       # quirk i1: curly quotes are not mapped
    assert norm.normalize("a, \"'\"), (b") == "a'b"                             # ... and this 9-character key is
    assert norm.normalize("好，，，好...") == "好,,,好…"                          # quirk i2
    assert fe.TextNormalizer().normalize("x") == ""                             # not loaded: the reference prints and returns ""
    with pytest.raises(ImportError, match="WeTextProcessing"):
        fe.TextNormalizer().load()


@needs_ref
def test_normaliser_equals_the_reference():
    ns = _reference_namespace()
    ref = _reference_normalizer(ns)
    got = fe.TextNormalizer(ZH, EN)
    assert got.char_rep_map == ref.char_rep_map and list(got.char_rep_map) == list(ref.char_rep_map)
    assert got.zh_char_rep_map == ref.zh_char_rep_map and list(got.zh_char_rep_map) == list(ref.zh_char_rep_map)
    for t in NORM_TEXTS:
        assert got.use_chinese(t) == ref.use_chinese(t), t
        assert got.normalize(t) == ref.normalize(t), t
    for p in ["xuan2", "ju4", "QUN1", "lüe4", "xue2", "ni3", "xu", "jü3", "quan12"]:
        assert got.correct_pinyin(p) == ref.correct_pinyin(p), p
    rng = random.Random(11)
    alphabet = list("你好吗克·—-xuanjqe12345 ,.:;!?…$()[]\"'，。：\n~ABz") + ["...", ",,,", "，，，", "……", "xuan2", "ju3", ", \"'\"), ("]
    for _ in range(500):
        t = "".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 30)))
        assert got.normalize(t) == ref.normalize(t), t


def _random_tokens(rng, n):
    alphabet = ["A", "B", "▁C", "▁", ",", "▁,", "-", ".", "▁.", "!", "?", "▁?", "▁...", "'", "▁'", "好"]
    weights = [8, 8, 8, 2, 3, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 6]
    return rng.choices(alphabet, weights, k=n)


def _run(split, tokens, marks, limit):
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        try:
            return [list(s) for s in split(list(tokens), marks, limit)], len(caught)
        except RecursionError:
            return "RecursionError", None


def test_sentence_splitter_fixed_cases():
    split = fe.TextTokenizer.split_sentences_by_token
    marks = fe.TextTokenizer.punctuation_marks_tokens
    assert split([], marks, 10) == []
    assert split(["A", ".", "B", "!", "C"], marks, 2) == [["A", "."], ["B", "!"], ["C"]]
    assert split(["A", ".", "B", "!", "C"], marks, 4) == [["A", ".", "B", "!"], ["C"]]          # neighbours glued while they fit
    assert split([".", "▁", ".", "A", "."], marks, 10) == [["A", "."]]                           # lone marks dropped
    assert split(["A", ".", "'", "B"], marks, 3) == [["A", ".", "'"], ["B"]]                     # a closing quote stays
    assert split(["A", ",", "B", ",", "C", "."], marks, 3) == [["A", ","], ["B", ","], ["C", "."]]
    assert split(["A", "-", "B", "-", "C", "."], marks, 3) == [["A", "-"], ["B", "-"], ["C", "."]]
    with pytest.warns(RuntimeWarning, match="exceeds limit"):
        assert split(["A", "B", "C", "D", "."], marks, 2) == [["A", "B"], ["C", "D", "."]]
    assert split(["A", "B", "C", "D"], marks, 2) == [["A", "B", "C", "D"]]                       # quirk i5: the tail is never cut
    with pytest.raises(RecursionError):                                                          # quirk i4
        split(["A", "B", "C", ",", "D", "."], marks, 3)


@needs_ref
def test_sentence_splitter_equals_the_reference():
    ref_split = _reference_namespace()["TextTokenizer"].split_sentences_by_token
    marks = fe.TextTokenizer.punctuation_marks_tokens
    rng = random.Random(23)
    depth = sys.getrecursionlimit()
    sys.setrecursionlimit(300)                           # the reference's endless recursion (i4) ends sooner
    try:
        outcomes = {"ok": 0, "warned": 0, "recursion": 0}
        for _ in range(3000):
            tokens = _random_tokens(rng, rng.randrange(0, 40))
            limit = rng.randrange(2, 14)
            want, got = _run(ref_split, tokens, marks, limit), _run(fe.TextTokenizer.split_sentences_by_token, tokens, marks, limit)
            assert got == want, (tokens, limit)
            outcomes["recursion" if want[0] == "RecursionError" else "warned" if want[1] else "ok"] += 1
        assert min(outcomes.values()) > 20, outcomes      # every branch was exercised
    finally:
        sys.setrecursionlimit(depth)


def test_tokenizer_fixed_cases(bpe_model):
    tok = fe.TextTokenizer(bpe_model, fe.TextNormalizer(ZH, EN))
    assert tok.bos_token_id == 0 and tok.eos_token_id == 1 and tok.pad_token_id == -1 and tok.unk_token_id == 2
    pieces = tok.tokenize("你好, hello world.")
    assert "".join(pieces).replace("▁", " ").strip() == "你 好 , HELLO WORLD."
    ids = tok.encode("你好, hello world.")
    assert ids == tok.convert_tokens_to_ids(pieces) and tok.convert_ids_to_tokens(ids) == pieces
    assert tok.decode(ids) == "你好,HELLO WORLD."
    assert tok.encode("") == [] and tok.tokenize("2") == tok.sp_model.Encode("2", out_type=str)    # quirk i3: "2" is not spoken
    assert tok.get_vocab()["<unk>"] == 2 and len(tok.get_vocab()) == tok.vocab_size
    with pytest.raises(ValueError, match="does not exist"):
        fe.TextTokenizer(bpe_model + ".missing")
    with pytest.raises(ValueError, match="None"):
        fe.TextTokenizer(None)


@needs_ref
def test_tokenizer_equals_the_reference(bpe_model):
    ns = _reference_namespace()
    ref = ns["TextTokenizer"](bpe_model, _reference_normalizer(ns))
    got = fe.TextTokenizer(bpe_model, fe.TextNormalizer(ZH, EN))
    assert got.vocab_size == ref.vocab_size and got.unk_token_id == ref.unk_token_id and got.special_tokens_map == ref.special_tokens_map
    assert got.punctuation_marks_tokens == ref.punctuation_marks_tokens
    texts = [t for t in NORM_TEXTS] + ["2", " 好 ", "大家好。我现在正在体验 ai 科技！你呢？Hello world, this is a test. I am here"]
    for t in texts:
        assert got.tokenize(t) == ref.tokenize(t), t
        assert got.encode(t) == ref.encode(t), t
        ids = ref.encode(t)
        assert got.decode(ids) == ref.decode(ids) and got.decode(ids, do_lower_case=True) == ref.decode(ids, do_lower_case=True)
        for limit in (4, 8, 120):
            want = _run(lambda *a: ref.split_sentences(a[0], max_tokens_per_sentence=a[2]), ref.tokenize(t), None, limit)
            have = _run(lambda *a: got.split_sentences(a[0], max_tokens_per_sentence=a[2]), got.tokenize(t), None, limit)
            assert have == want, (t, limit)
    assert got.batch_encode(texts[:5]) == ref.batch_encode(texts[:5])
    assert got.convert_tokens_to_ids("▁") == ref.convert_tokens_to_ids("▁")


class _FakeEngine:
    """Records what the sentence loop asks of the engine; the hidden rows encode the sentence so the output can be traced."""
    MEL_CODES = 16

    def __init__(self):
        self.calls = []

    def indextts_gpt_info(self):
        return {"dim": 4, "layers": 1, "heads": 1, "mel_codes": self.MEL_CODES, "max_rows": 64}

    def indextts_gpt_generate(self, conds_latent, text_ids, max_new=0, precision=None, penalty=None):
        k = sum(c[0] == "gpt" for c in self.calls)
        self.calls.append(("gpt", np.array(text_ids), penalty.copy(), max_new))
        n = 3 + k
        out = penalty.copy()
        out[0, k] = 0.7                                   # the engine hands the updated penalty back
        return np.arange(n, dtype=np.int32), np.full((n, 4), float(k + 1), np.float32), out

    def indextts_vocoder_run(self, hidden, conds, cond_layer, precision=None):
        self.calls.append(("vocoder", hidden.copy(), len(conds), np.asarray(cond_layer).copy()))
        return np.full((1, 1, 1024 * (hidden.shape[0] - 2) + 30), int(hidden[0, 0]), dtype=np.int16)


def test_sentence_loop_on_a_recording_engine(bpe_model, tmp_path):
    tok = fe.TextTokenizer(bpe_model, fe.TextNormalizer(ZH, EN))
    eng = _FakeEngine()
    syn = fe.IndexTTSSynthesizer(eng, tok, precision=1, max_tokens_per_sentence=6)
    text = "大家好。hello world! 你好世界?"
    sents = syn.sentences(text)
    assert len(sents) == 3 and sents[0][0].replace(" ", "") == "大家好."
    conditioning = [np.full((1, 8, 1), i, np.float32) for i in range(6)] + [np.full((1, 8, 1), 9, np.float32), np.zeros((1, 32, 4), np.float32)]
    out = tmp_path / "generated.wav"
    wav = syn.synthesize(conditioning, text, out_path=str(out))
    gpt_calls = [c for c in eng.calls if c[0] == "gpt"]
    voc_calls = [c for c in eng.calls if c[0] == "vocoder"]
    assert [list(c[1][0]) for c in gpt_calls] == [s[2] for s in sents]
    assert gpt_calls[0][2].sum() == eng.MEL_CODES                                     # all ones into the first sentence ...
    assert gpt_calls[2][2][0, 0] == np.float32(0.7) and gpt_calls[2][2][0, 1] == np.float32(0.7)   # ... carried afterwards (i7)
    assert all(c[2] == 6 and float(c[3].ravel()[0]) == 9.0 for c in voc_calls)
    pad = int(24000 * 0.2)
    last = 1024 * (5 - 2) + 30
    assert wav.shape == (1, 1, last + pad) and wav.dtype == np.int16                  # the reference's file: the LAST sentence (i7)
    assert (wav[0, 0, :last] == 3).all() and (wav[0, 0, last:] == 0).all()
    np.testing.assert_array_equal(fe.load_wav_mono_int16(str(out)), wav.reshape(-1))
    eng2 = _FakeEngine()
    whole = fe.IndexTTSSynthesizer(eng2, tok, precision=1, max_tokens_per_sentence=6).synthesize(conditioning, text, keep="all")
    assert whole.shape[-1] == sum(1024 * (n - 2) + 30 + pad for n in (3, 4, 5))
    with pytest.raises(ValueError, match="no sentence"):
        syn.synthesize(conditioning, "")
    with pytest.raises(ValueError, match="keep"):
        syn.synthesize(conditioning, text, keep="first")
