"""Host front end of the IndexTTS call surface: text in, sentences of BPE ids out, and the per-sentence loop around the engine.

Mirrors what the reference's driver script does around its `InferenceSession.run` calls (IndexTTS/Inference_IndexTTS_ONNX.py):
CJK pre-tokenisation (:94-125) and its inverse (:128-170), the text normaliser's language switch, pinyin / name protection and
punctuation table (:173-356), the SentencePiece tokenizer and the sentence splitter (:359-574), and the sentence loop (:719-804).
Class and method names are the reference's, so a script written against it keeps working. Pure host Python: nothing here is on
the measured hot path.

Not installable offline: `tn` (WeTextProcessing) / `wetext`, the number / date verbaliser the reference loads in
`TextNormalizer.load`. It is used when present; without it `load()` raises unless the caller passes its own
`zh_normalizer` / `en_normalizer` objects (anything with `.normalize(str) -> str`; `IdentityNormalizer` leaves digits unspoken).

Quirks of the reference that are reproduced because results have to be identical (each one is tested against the reference's
own code, AST-extracted where it lies):
  i1  the punctuation table's two entries for the curly double quotes were flattened to ASCII quotes in the script, so three
      ASCII double quotes in a row open a triple-quoted string there and the two entries collapse into ONE: the 9-character
      key  , DQUOTE ' DQUOTE ) , SPACE (  maps to ', and curly quotes are NOT mapped at all (:185; `_PUNCT_TABLE` below).
  i2  the table is applied as one regex alternation in table order, so the 3-character keys `，，，` is dead (the single `，`
      wins first) while `...`, `,,,` and `……` are live (:206-207).
  i3  `encode` skips normalisation for a text that strips to ONE character (:466-468).
  i4  a sentence longer than the limit whose only comma (or hyphen) is its last token recurses without end in the reference
      (:538-541 call :487 with the same arguments); here the same case raises RecursionError at once.
  i5  the trailing tokens after the last sentence mark are kept as a sentence without a length check (:531-532).
  i6  `de_tokenized_by_CJK_char` restores only the first English placeholder of every blank-separated word (:160-168).
  i7  the script writes only the LAST sentence's waveform (`generated_wav` is overwritten per sentence and `save_generated_wav`
      is never filled, :724 / :793 / :804), and the repeat-penalty vector is carried from one sentence to the next (:771-774).
      `IndexTTSSynthesizer.synthesize(..., keep="last")` is the reference's file; `keep="all"` concatenates the sentences.
"""
import os
import re
import warnings

import numpy as np

from .frontend import load_wav_mono_int16, save_wav  # noqa: F401  (same wav conventions: mono int16 at 24 kHz)

SAMPLE_RATE = 24000
STOP_TOKEN = (8193,)
MAX_GENERATE_LENGTH = 800
SPLIT_PAD_SECONDS = 0.2

# Hangul Jamo, CJK radicals .. Yi, Phags-pa .. Hangul syllables, CJK compatibility ideographs / forms, half-width kana and
# Hangul, the supplementary ideographic plane (:113)
_CJK_CHAR = re.compile("([ᄀ-ᇿ⺀-꓏ꡀ-힯豈-﫿︰-﹏･-ￜ\U00020000-\U0002ffff])")
_ENGLISH_RUN = re.compile(r"([A-Z]+(?:[\s-][A-Z-]+)*)", re.IGNORECASE)
_PLACEHOLDER = re.compile(r"^.*?(<sent_(\d+)>)")


def tokenize_by_CJK_char(line: str, do_upper_case=True) -> str:
    """"你好世界是 hello world 的中文" -> "你 好 世 界 是 HELLO WORLD 的 中 文": every CJK character becomes its own blank-separated
    word, everything else is upper-cased (:94-125)."""
    pieces = (p.strip() for p in _CJK_CHAR.split(line.strip()))
    kept = [p for p in pieces if p]
    return " ".join(p.upper() for p in kept) if do_upper_case else " ".join(kept)


def de_tokenized_by_CJK_char(line: str, do_lower_case=False) -> str:
    """The inverse: blanks between CJK characters vanish, blanks inside English runs stay (:128-170, quirk i6)."""
    runs = _ENGLISH_RUN.findall(line)
    if not runs:
        return "".join(line.split())
    masked = line
    for i, run in enumerate(runs):                       # first occurrence each, in order: later runs may land in earlier tags
        masked = masked.replace(run, "<sent_%d>" % i, 1)
    words = masked.split()
    for j, word in enumerate(words):
        m = _PLACEHOLDER.match(word)
        if m is None:
            continue
        word = word.replace(m.group(1), runs[int(m.group(2))])
        words[j] = word.lower() if do_lower_case else word
    return "".join(words)


# (key, replacement) in application order; see i1 / i2
_PUNCT_TABLE = (
    ("：", ","), ("；", ","), (";", ","), ("，", ","), ("。", "."), ("！", "!"), ("？", "?"), ("\n", " "), ("·", "-"), ("、", ","),
    ("...", "…"), (",,,", "…"), ("，，，", "…"), ("……", "…"),
    (", \"'\"), (", "'"),                                # i1
    ('"', "'"), ("'", "'"),
    ("（", "'"), ("）", "'"), ("(", "'"), (")", "'"), ("《", "'"), ("》", "'"), ("【", "'"), ("】", "'"), ("[", "'"), ("]", "'"),
    ("—", "-"), ("～", "-"), ("~", "-"), ("「", "'"), ("」", "'"), (":", ","),
)


def _table_substituter(table):
    mapping = dict(table)
    pattern = re.compile("|".join(re.escape(k) for k in mapping))
    return lambda text: pattern.sub(lambda m: mapping[m.group()], text)


class IdentityNormalizer:
    """Stand-in verbaliser: returns the text unchanged (numbers, dates and units stay as written)."""

    def normalize(self, text):
        return text


class TextNormalizer:
    """Language switch + protection of inline pinyin / hyphenated names around the external verbaliser + punctuation table."""

    _EMAIL = re.compile(r"^[a-zA-Z0-9]+@[a-zA-Z0-9]+\.[a-zA-Z]+$")
    _HAN = re.compile("[一-鿿]")
    _LATIN = re.compile(r"[a-zA-Z]")
    _PINYIN_TONE = re.compile(r"([bmnpqdfghjklzcsxwy]?h?[aeiouüv]{1,2}[ng]*|ng)([1-5])", re.IGNORECASE)
    _NAME = re.compile("[一-鿿]+([-·—][一-鿿]+){1,2}")
    _JQX_U = re.compile(r"([jqx])[uü](n|e|an)*(\d)", re.IGNORECASE)

    def __init__(self, zh_normalizer=None, en_normalizer=None):
        self.zh_normalizer = zh_normalizer
        self.en_normalizer = en_normalizer
        self.char_rep_map = dict(_PUNCT_TABLE)
        self.zh_char_rep_map = {"$": ".", **self.char_rep_map}
        self._sub_en = _table_substituter(_PUNCT_TABLE)
        self._sub_zh = _table_substituter((("$", "."),) + _PUNCT_TABLE)

    def load(self):
        """The reference's `load` (:225-234) picks WeTextProcessing (`tn`) on Linux / Windows and `wetext` on macOS, with the same
        constructor arguments as here; this one takes whichever of the two imports, `tn` first, on any platform. Normalisers passed
        to the constructor are kept."""
        if self.zh_normalizer is not None and self.en_normalizer is not None:
            return
        try:
            from tn.chinese.normalizer import Normalizer as Zh
            from tn.english.normalizer import Normalizer as En
            zh, en = Zh(remove_interjections=False, remove_erhua=False, overwrite_cache=False), En(overwrite_cache=False)
        except ImportError:
            try:
                from wetext import Normalizer
                zh, en = Normalizer(remove_erhua=False, lang="zh", operator="tn"), Normalizer(lang="en", operator="tn")
            except ImportError as exc:
                raise ImportError("IndexTTS text normalisation needs WeTextProcessing (`tn`) or `wetext` (neither is installed): "
                                  "pass zh_normalizer= / en_normalizer= (e.g. IdentityNormalizer()) to TextNormalizer") from exc
        self.zh_normalizer = self.zh_normalizer or zh
        self.en_normalizer = self.en_normalizer or en

    def match_email(self, email):
        return self._EMAIL.match(email) is not None

    def use_chinese(self, s):
        """Chinese rules when there is a Han character, no Latin letter at all, a bare e-mail address, or inline pinyin (:212-223)."""
        if self._HAN.search(s) or not self._LATIN.search(s) or self.match_email(s):
            return True
        return self._PINYIN_TONE.search(s) is not None

    def normalize(self, text: str) -> str:
        text = text.replace("嗯", "恩").replace("呣", "母")
        if not self.zh_normalizer or not self.en_normalizer:
            print("Error, text normalizer is not initialized !!!")
            return ""
        text = text.rstrip()
        if not self.use_chinese(text):
            try:
                spoken = self.en_normalizer.normalize(text)
            except Exception:                            # the reference prints the traceback and keeps the raw text (:262-265)
                import traceback
                print(traceback.format_exc())
                spoken = text
            return self._sub_en(spoken)
        guarded, pinyins = self.save_pinyin_tones(text)
        guarded, names = self.save_names(guarded)
        try:
            spoken = self.zh_normalizer.normalize(guarded)
        except Exception:                                # ... and an empty sentence on the Chinese side (:251-254)
            import traceback
            print(traceback.format_exc())
            spoken = ""
        spoken = self.restore_pinyin_tones(self.restore_names(spoken, names), pinyins)
        return self._sub_zh(spoken)

    def correct_pinyin(self, pinyin: str):
        """ju / qu / xu (+ n, e, an) are written with v, upper-cased: "xuan2" -> "XVAN2"; anything else is untouched (:270-277)."""
        if pinyin[0].lower() not in "jqx":
            return pinyin
        return self._JQX_U.sub(r"\g<1>v\g<2>\g<3>", pinyin).upper()

    @staticmethod
    def _guard(text, found, tag):
        """Replace every distinct match (first-seen order) by <tag_a>, <tag_b>, ... -> (text, matches or None)."""
        distinct = list(dict.fromkeys("".join(groups) for groups in found))
        if not distinct:
            return text, None
        for i, item in enumerate(distinct):
            text = text.replace(item, "<%s_%s>" % (tag, chr(ord("a") + i)))
        return text, distinct

    @staticmethod
    def _unguard(text, saved, tag, fix=lambda s: s):
        for i, item in enumerate(saved or ()):
            text = text.replace("<%s_%s>" % (tag, chr(ord("a") + i)), fix(item))
        return text

    def save_names(self, original_text):
        # findall yields only the LAST "-part" group of each name (the pattern has one group), so that is what gets guarded (:281-303)
        return self._guard(original_text, self._NAME.findall(original_text), "n")

    def restore_names(self, normalized_text, original_name_list):
        return self._unguard(normalized_text, original_name_list, "n")

    def save_pinyin_tones(self, original_text):
        return self._guard(original_text, self._PINYIN_TONE.findall(original_text), "pinyin")

    def restore_pinyin_tones(self, normalized_text, original_pinyin_list):
        return self._unguard(normalized_text, original_pinyin_list, "pinyin", self.correct_pinyin)


class TextTokenizer:
    """SentencePiece ids after normalisation and CJK pre-tokenisation, and the sentence splitter (:359-574)."""

    punctuation_marks_tokens = [".", "!", "?", "▁.", "▁?", "▁..."]
    unk_token, pad_token, bos_token, eos_token = "<unk>", None, "<s>", "</s>"
    pad_token_id, bos_token_id, eos_token_id = -1, 0, 1

    def __init__(self, vocab_file: str, normalizer: TextNormalizer = None):
        if vocab_file is None:
            raise ValueError("vocab_file is None")
        if not os.path.exists(vocab_file):
            raise ValueError(f"vocab_file {vocab_file} does not exist")
        from sentencepiece import SentencePieceProcessor
        self.vocab_file = vocab_file
        self.normalizer = normalizer
        if normalizer:
            normalizer.load()
        self.sp_model = SentencePieceProcessor(model_file=vocab_file)
        self.pre_tokenizers = [tokenize_by_CJK_char]
        self._vocab = None

    @property
    def vocab_size(self):
        return self.sp_model.GetPieceSize()

    @property
    def unk_token_id(self):
        return self.sp_model.unk_id()

    @property
    def special_tokens_map(self):
        return {"unk_token": self.unk_token, "pad_token": self.pad_token, "bos_token": self.bos_token, "eos_token": self.eos_token}

    def get_vocab(self):
        if self._vocab is None:
            self._vocab = {self.sp_model.IdToPiece(i): i for i in range(self.vocab_size)}
        return self._vocab

    def convert_ids_to_tokens(self, ids):
        return self.sp_model.IdToPiece(ids)

    def convert_tokens_to_ids(self, tokens):
        return [self.sp_model.PieceToId(t) for t in ([tokens] if isinstance(tokens, str) else tokens)]

    def _prepare(self, text):
        if self.normalizer:
            text = self.normalizer.normalize(text)
        for pre in self.pre_tokenizers:
            text = pre(text)
        return text

    def encode(self, text: str, **kwargs):
        if not text:
            return []
        out_type = kwargs.pop("out_type", int)
        if len(text.strip()) == 1:                       # i3
            return self.sp_model.Encode(text, out_type=out_type, **kwargs)
        return self.sp_model.Encode(self._prepare(text), out_type=out_type, **kwargs)

    def tokenize(self, text: str):
        return self.encode(text, out_type=str)

    def batch_encode(self, texts, **kwargs):
        return self.sp_model.Encode([self._prepare(t) for t in texts], out_type=kwargs.pop("out_type", int), **kwargs)

    def decode(self, ids, do_lower_case=False, **kwargs):
        ids = [ids] if isinstance(ids, int) else ids
        return de_tokenized_by_CJK_char(self.sp_model.Decode(ids, out_type=kwargs.pop("out_type", str), **kwargs), do_lower_case=do_lower_case)

    @staticmethod
    def split_sentences_by_token(tokenized_str, split_tokens, max_tokens_per_sentence):
        """Cut after every split token (a following quote token stays with the sentence), drop chunks that are a lone mark, cut
        oversized chunks at commas, else hyphens, else in two at the limit, then glue neighbours while they fit (:486-565)."""
        marks = set(split_tokens)
        limit = max_tokens_per_sentence
        chunks, chunk = [], []
        n = len(tokenized_str)
        i = 0
        while i < n:
            tok = tokenized_str[i]
            i += 1
            chunk.append(tok)
            if tok not in marks:
                continue
            if len(chunk) <= 1 or (len(chunk) == 2 and chunk[0] == "▁"):
                chunk = []
                continue
            if i < n and tokenized_str[i] in ("'", "▁'"):
                chunk.append(tokenized_str[i])
                i += 1
            if len(chunk) <= limit:
                chunks.append(chunk)
            else:
                finer = TextTokenizer._finer_marks(chunk)
                if finer is None:
                    warnings.warn(f"The tokens length of sentence exceeds limit: {limit}, Tokens in sentence: {chunk}. "
                                  "Maybe unexpected behavior", RuntimeWarning)
                    chunks += [chunk[:limit], chunk[limit:]]
                elif len(chunk) == n and set(finer) == marks:                                          # i4
                    raise RecursionError("sentence splitter: an oversized sentence whose only %r is its last token cannot be "
                                         "split (the reference recurses without end here)" % (finer[0],))
                else:
                    chunks += TextTokenizer.split_sentences_by_token(chunk, finer, limit)
            chunk = []
        if chunk:                                        # i5
            chunks.append(chunk)
        merged = []
        for c in chunks:
            if not c:
                continue
            if merged and len(merged[-1]) + len(c) <= limit:
                merged[-1].extend(c)
            else:
                merged.append(c)
        return merged

    @staticmethod
    def _finer_marks(sentence):
        if "," in sentence or "▁," in sentence:
            return [",", "▁,"]
        if "-" in sentence:
            return ["-"]
        return None

    def split_sentences(self, tokenized, max_tokens_per_sentence=120):
        return self.split_sentences_by_token(tokenized, self.punctuation_marks_tokens, max_tokens_per_sentence)


class IndexTTSSynthesizer:
    """The reference script's sentence loop (:719-804) on the engine: text -> sentences -> per sentence the GPT-2 greedy decode
    (graphs B, C, D, E fused into one device loop) and the IndexTTS_F vocoder -> int16 waveform + 200 ms of silence.

    engine: a capi.Engine with the "igpt." / "ivgan." tensors loaded and both halves built.
    conditioning: the outputs of the reference's conditioning graph IndexTTS_A in ITS output order (Export_IndexTTS.py:200):
    `[*save_bigvgan_conds, bigvgan_cond_layer_speaker_embedding, conds_latent]`. Graph A itself is not part of this engine
    (DESIGN.md 7); its outputs depend only on the reference voice, so they are computed once per voice wherever graph A runs.
    """

    def __init__(self, engine, tokenizer, precision=None, max_tokens_per_sentence=120, sample_rate=SAMPLE_RATE):
        from . import capi
        self.engine = engine
        self.tokenizer = tokenizer
        self.precision = capi.BF16 if precision is None else precision
        self.max_tokens_per_sentence = max_tokens_per_sentence
        self.sample_rate = sample_rate
        self.split_pad = np.zeros((1, 1, int(sample_rate * SPLIT_PAD_SECONDS)), dtype=np.int16)
        self.trace = []

    def sentences(self, gen_text):
        """-> [(display text, [token], [id])] the way the script prints and feeds them (:721-732)."""
        pieces = self.tokenizer.split_sentences(self.tokenizer.tokenize(gen_text), self.max_tokens_per_sentence)
        return [("".join(p).replace("▁", " "), p, self.tokenizer.convert_tokens_to_ids(p)) for p in pieces]

    def synthesize(self, conditioning, gen_text, out_path=None, keep="last", max_new=0):
        if keep not in ("last", "all"):
            raise ValueError("keep must be 'last' (the reference's output file, quirk i7) or 'all'")
        conditioning = list(conditioning)
        vocoder_conds, cond_layer, conds_latent = conditioning[:-2], conditioning[-2], conditioning[-1]
        mel_codes = self.engine.indextts_gpt_info()["mel_codes"]
        penalty = np.ones((1, mel_codes), dtype=np.float32)      # carried across sentences (i7)
        waves = []
        self.trace = []                                  # per sentence: what the script prints (:728, :782) and the mel ids
        for shown, _, ids in self.sentences(gen_text):
            mel_ids, hidden, penalty = self.engine.indextts_gpt_generate(conds_latent, np.asarray([ids], dtype=np.int32), max_new=max_new,
                                                                   precision=self.precision, penalty=penalty)
            pcm = self.engine.indextts_vocoder_run(hidden, vocoder_conds, cond_layer, precision=self.precision)
            waves.append(np.concatenate([pcm, self.split_pad], axis=-1))
            self.trace.append({"text": shown, "text_ids": list(ids), "mel_ids": np.asarray(mel_ids).copy(), "samples": int(pcm.shape[-1])})
        if not waves:                                    # the script would hit an undefined `generated_wav` (:804)
            raise ValueError("no sentence to speak in %r" % (gen_text,))
        wav = waves[-1] if keep == "last" else np.concatenate(waves, axis=-1)
        if out_path is not None:
            save_wav(out_path, wav, self.sample_rate)
        return wav
