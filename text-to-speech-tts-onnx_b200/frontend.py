"""Host front end of the F5-TTS call surface: reference wav + reference text + text to speak in, waveform out.

Mirrors what the reference's driver script does around its three `InferenceSession.run` calls
(F5_TTS/F5-TTS-ONNX-Inference.py): vocabulary (:86-90), text -> symbols (:96-136), symbols -> ids (:140-148), audio loading
(:223-225), the duration heuristic (:227-231, quirk q13 included) and wav writing (:315). Pure host Python: nothing here is on
the measured hot path. `jieba` / `pypinyin` (Chinese word segmentation / pinyin) and `pydub` / `soundfile` are not installable
offline; they are used when present and replaced as documented below when absent.
"""
import os
import re
import wave

import numpy as np

SAMPLE_RATE = 24000
HOP_LENGTH = 256

# the reference maps these before looking symbols up (":" / quotes that are out of vocabulary, :104-106)
_PUNCT_MAP = str.maketrans({";": ",", "“": '"', "”": '"', "‘": "'", "’": "'"})
# the reference's pause-punctuation "class" is written as a plain string, so re.findall looks for these seven characters IN A
# ROW (quirk q13): the "+3 per pause mark" term is zero for any normal sentence. Reproduced as is.
_ZH_PAUSE_PATTERN = r"。，、；：？！"


def load_vocab(path):
    """vocab.txt: one symbol per line, index = line number. The reference drops the LAST character of every line (the newline,
    :89), so a final line without a newline loses its last character there too -- reproduced."""
    table = {}
    with open(path, "r", encoding="utf-8") as f:
        for i, line in enumerate(f):
            table[line[:-1]] = i
    return table


def _is_cjk(ch):
    return "㄀" <= ch <= "鿿"


def _fallback_segments(text):
    """Stand-in for jieba.cut when jieba is absent: runs of ASCII letters / digits / apostrophes are one segment, every other
    character is its own segment. For ASCII text this is what jieba's default mode yields (its regex splits on non-word
    characters); Chinese text needs jieba for word boundaries, which only matter to pypinyin's tone sandhi."""
    return [m.group(0) for m in re.finditer(r"[A-Za-z0-9']+|.", text, flags=re.S)]


def text_to_symbols(texts, polyphone=True, segmenter=None, to_pinyin=None):
    """[str] -> [[symbol]] the way convert_char_to_pinyin does (F5-TTS-ONNX-Inference.py:96-136).

    segmenter(text) -> iterable of segments (default: jieba.cut, else `_fallback_segments`);
    to_pinyin(str) -> list of TONE3 syllables, one per character (default: pypinyin.lazy_pinyin(.., Style.TONE3, tone_sandhi=True)).
    """
    if segmenter is None:
        try:
            import jieba
            if not jieba.dt.initialized:
                jieba.default_logger.setLevel(50)
                jieba.initialize()
            segmenter = jieba.cut
        except ImportError:
            segmenter = _fallback_segments
    if to_pinyin is None:
        try:
            from pypinyin import Style, lazy_pinyin

            def to_pinyin(s):
                return lazy_pinyin(s, style=Style.TONE3, tone_sandhi=True)
        except ImportError:
            def to_pinyin(s):
                raise ImportError("Chinese text needs pypinyin (not installed): pass to_pinyin= or install it")
    out = []
    for text in texts:
        symbols = []
        for seg in segmenter(text.translate(_PUNCT_MAP)):
            nbytes = len(seg.encode("utf-8"))
            if nbytes == len(seg):                       # ASCII only: letters, digits, symbols
                if symbols and nbytes > 1 and symbols[-1] not in " :'\"":
                    symbols.append(" ")                  # a word glued to what precedes it gets a separating space
                symbols.extend(seg)
            elif polyphone and nbytes == 3 * len(seg):   # only 3-byte (east asian) characters: pinyin of the whole word
                syl = to_pinyin(seg)
                for ch, s in zip(seg, syl):
                    if _is_cjk(ch):
                        symbols.append(" ")
                    symbols.append(s)
            else:                                        # mixed segment: character by character
                for ch in seg:
                    if ord(ch) < 256:
                        symbols.append(ch)
                    elif _is_cjk(ch):
                        symbols.append(" ")
                        symbols.extend(to_pinyin(ch))
                    else:
                        symbols.append(ch)
        out.append(symbols)
    return out


def symbols_to_ids(symbol_lists, vocab, padding_value=-1):
    """[[symbol]] -> int32 (batch, longest), unknown symbols -> 0, short rows padded (list_str_to_idx, :140-148)."""
    rows = [[vocab.get(s, 0) for s in syms] for syms in symbol_lists]
    width = max((len(r) for r in rows), default=0)
    ids = np.full((len(rows), width), padding_value, dtype=np.int32)
    for i, r in enumerate(rows):
        ids[i, :len(r)] = r
    return ids


def estimate_max_duration(ref_text, gen_text, audio_len, speed=1.0, hop=HOP_LENGTH):
    """max_duration (frames) = reference frames + reference frames * gen / ref text length / speed (:227-231). Text length is the
    UTF-8 byte count plus 3 per match of the pause pattern (q13: a literal 7-character string)."""
    ref_len = len(ref_text.encode("utf-8")) + 3 * len(re.findall(_ZH_PAUSE_PATTERN, ref_text))
    gen_len = len(gen_text.encode("utf-8")) + 3 * len(re.findall(_ZH_PAUSE_PATTERN, gen_text))
    frames = audio_len // hop + 1
    return frames + int(frames / ref_len * gen_len / speed)


def load_wav_mono_int16(path, sample_rate=SAMPLE_RATE):
    """PCM wav -> mono int16 at `sample_rate` (the reference: pydub AudioSegment.set_channels(1).set_frame_rate(24000), :223).
    Channels are averaged as pydub / audioop.tomono do; other sample rates are resampled by linear interpolation (audioop.ratecv
    is a linear interpolator too, with a different phase: bit-equality with pydub holds for 24 kHz input only)."""
    with wave.open(path, "rb") as w:
        nch, width, rate, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = w.readframes(n)
    if width == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float64)
    elif width == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float64) - 128.0) * 256.0
    elif width == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float64) / 65536.0
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        x = (np.where(v >= 1 << 23, v - (1 << 24), v) / 256.0).astype(np.float64)
    else:
        raise ValueError(f"{path}: unsupported sample width {width}")
    x = x.reshape(-1, nch)
    mono = np.floor(x.mean(axis=1)) if nch > 1 else x[:, 0]          # audioop.tomono: floor(0.5 l + 0.5 r)
    if rate != sample_rate and mono.size:
        n_out = int(round(mono.size * sample_rate / rate))
        mono = np.floor(np.interp(np.arange(n_out) * (rate / sample_rate), np.arange(mono.size), mono) + 0.5)
    return np.clip(mono, -32768, 32767).astype(np.int16)


def save_wav(path, pcm, sample_rate=SAMPLE_RATE):
    """int16 PCM -> wav (the reference writes WAVE_FORMAT_EXTENSIBLE through soundfile, :315; plain PCM carries the same samples)."""
    pcm = np.ascontiguousarray(np.asarray(pcm).reshape(-1), dtype="<i2")
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(int(sample_rate))
        w.writeframes(pcm.tobytes())


class F5Synthesizer:
    """`synthesize(reference_audio, ref_text, gen_text) -> int16 waveform`: the reference script's whole flow on the engine.

    engine: a capi.Engine with the F5 tensors loaded and built (and BigVGAN's, for vocoder="bigvgan").
    The Euler start noise is drawn here from numpy's generator (seed 9527, the reference's RANDOM_SEED): the reference draws it
    inside graph A with ORT's RandomNormalLike, whose stream cannot be reproduced outside ORT ("Noise").
    """

    def __init__(self, engine, vocab, precision=None, nfe_steps=-1, speed=1.0, seed=9527):
        from . import capi
        self.engine = engine
        self.vocab = load_vocab(vocab) if isinstance(vocab, (str, os.PathLike)) else dict(vocab)
        self.precision = capi.F16 if precision is None else precision
        self.nfe_steps, self.speed, self.seed = nfe_steps, speed, seed

    def prepare(self, reference_audio, ref_text, gen_text):
        """-> (audio int16 (1,1,L), text_ids int32 (1,n), max_duration, noise (1,N,100)): the inputs of graph A / B."""
        audio = load_wav_mono_int16(reference_audio) if isinstance(reference_audio, (str, os.PathLike)) else np.asarray(reference_audio, np.int16)
        audio = audio.reshape(-1)
        max_duration = estimate_max_duration(ref_text, gen_text, audio.size, self.speed)
        ids = symbols_to_ids(text_to_symbols([ref_text + gen_text]), self.vocab)
        noise = np.random.default_rng(self.seed).standard_normal((1, max_duration, 100), dtype=np.float32)
        return audio.reshape(1, 1, -1), ids, max_duration, noise

    def synthesize(self, reference_audio, ref_text, gen_text, out_path=None, vocoder="vocos"):
        audio, ids, max_duration, noise = self.prepare(reference_audio, ref_text, gen_text)
        if vocoder == "vocos":                                  # the reference's own graph C (Vocos + ISTFT)
            pcm = self.engine.f5_synthesize(audio, ids, max_duration, noise, precision=self.precision, n_steps=self.nfe_steps)
        elif vocoder == "bigvgan":                              # BASELINE.json's pipeline: the generated frames through BigVGAN
            pcm = self.engine.f5_bigvgan_pipeline(audio.reshape(1, -1), ids, max_duration, noise, precision=self.precision,
                                                  n_steps=self.nfe_steps)
        else:
            raise ValueError("vocoder must be 'vocos' or 'bigvgan'")
        pcm = np.asarray(pcm).reshape(-1)
        if out_path is not None:
            save_wav(out_path, pcm)
        return pcm
