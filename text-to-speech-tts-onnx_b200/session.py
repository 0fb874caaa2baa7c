"""Drop-in for the subset of ``onnxruntime`` the reference's drivers use on the hot path.

    import b200tts.session as onnxruntime          # the only line a reference script changes

The reference builds one ``InferenceSession`` per exported graph and calls ``.run`` / ``.run_with_ort_values``
(F5_TTS/F5-TTS-ONNX-Inference.py:173-311, BigVGAN/Export_BigVGAN.py:153-177). Here the graph is identified by
the file name of the model path (``BigVGAN.onnx``, ``F5_Preprocess.onnx``, ``F5_Transformer.onnx``,
``F5_Decode.onnx``) and executed by libb200tts on the GPU; input / output names, dtypes, shapes and ordering are
the reference's (Export_F5.py:294-306,354-365,409-414; Export_BigVGAN.py:65-70). Providers and session options
are accepted and ignored. Weights come from ``register_checkpoint`` (the engine has no ``.onnx`` initializers).
"""
import os

import numpy as np

from . import capi, weights
from .config import BIGVGAN, F5

_engines = {}
_checkpoints = {}
_seed = [None]


def get_engine(device: int = 0) -> capi.Engine:
    if device not in _engines:
        _engines[device] = capi.Engine(device)
    return _engines[device]


def register_checkpoint(kind: str, state: dict):
    """kind in {'bigvgan', 'dit', 'vocos'}; state = the reference's state dict for that model."""
    _checkpoints[kind] = state


def set_seed(seed: int):
    """onnxruntime.set_seed (F5-TTS-ONNX-Inference.py:152): seeds the Euler start noise of F5_Preprocess."""
    _seed[0] = int(seed)


# -- inert option holders ------------------------------------------------------------------------
class SessionOptions:
    def __init__(self):
        self._entries = {}

    def add_session_config_entry(self, k, v):
        self._entries[k] = v


class GraphOptimizationLevel:
    ORT_DISABLE_ALL, ORT_ENABLE_BASIC, ORT_ENABLE_EXTENDED, ORT_ENABLE_ALL = range(4)


class ExecutionMode:
    ORT_SEQUENTIAL, ORT_PARALLEL = range(2)


class NodeArg:
    def __init__(self, name, type_, shape):
        self.name, self.type, self.shape = name, type_, shape


class OrtValue:
    """Host-side value holder with the two static helpers the reference uses."""

    def __init__(self, array):
        self._array = array

    @staticmethod
    def ortvalue_from_numpy(array, device_type="cpu", device_id=0):
        return OrtValue(np.ascontiguousarray(array))

    def numpy(self):
        return self._array

    def shape(self):
        return list(self._array.shape)


class IOBinding:
    """The subset of onnxruntime's SessionIOBinding the reference's (dormant) bound loop uses
    (F5_TTS/F5-TTS-ONNX-Inference.py:257-288): inputs and outputs are OrtValues bound by name, outputs may alias inputs
    (`noise`, `time_step` are updated in place by every run_with_iobinding call)."""

    def __init__(self, session):
        self._session = session
        self._inputs, self._outputs = {}, {}

    def bind_ortvalue_input(self, name, ortvalue):
        self._inputs[name] = ortvalue

    def bind_ortvalue_output(self, name, ortvalue):
        self._outputs[name] = ortvalue

    def get_outputs(self):
        return [self._outputs[o.name] for o in self._session.get_outputs() if o.name in self._outputs]

    def clear_binding_inputs(self):
        self._inputs.clear()

    def clear_binding_outputs(self):
        self._outputs.clear()


def _as_numpy(v):
    return v.numpy() if isinstance(v, OrtValue) else np.asarray(v)


class _Graph:
    inputs = ()
    outputs = ()

    def run(self, feed: dict) -> list:
        raise NotImplementedError


def _has_state(engine, *prefixes):
    """Whether the engine already holds these weight prefixes (capi.Engine.load_state / load_blob); stand-in engines may not know."""
    f = getattr(engine, "has_state", None)
    return bool(f and f(*prefixes))


class _BigVGANGraph(_Graph):
    """BigVGAN/Export_BigVGAN.py:37-49,65-70."""

    def __init__(self, engine, precision, state=None):
        self.engine, self.precision = engine, precision
        self.cfg = BIGVGAN
        self.inputs = (NodeArg("mel_features", "tensor(float16)" if precision == capi.F16 else "tensor(float)",
                               [1, self.cfg.num_mels, "mel_features_len"]),)
        self.outputs = (NodeArg("generated_wav", "tensor(int16)", [1, 1, "generated_len"]),)
        state = state if state is not None else _checkpoints.get("bigvgan")
        if state is None and _has_state(engine, "bigvgan"):
            engine.bigvgan_build()           # weights came from an engine blob (Engine.load_blob): nothing to convert or upload
            return
        if state is None:
            raise RuntimeError("no BigVGAN checkpoint registered: call session.register_checkpoint('bigvgan', state) or engine.load_blob(path)")
        engine.load_state("bigvgan", weights.bigvgan_engine_tensors(state))
        engine.bigvgan_build()

    def run(self, feed):
        mel = _as_numpy(feed["mel_features"]).astype(np.float32)     # fp16 graphs feed float16
        if mel.ndim != 3 or mel.shape[1] != self.cfg.num_mels or mel.shape[2] == 0:
            raise ValueError(f"mel_features must be (B, {self.cfg.num_mels}, T>0), got {mel.shape}")
        return [self.engine.bigvgan_run(mel, precision=self.precision, hop=self.cfg.hop)]


class _IndexTTSVocoderGraph(_Graph):
    """IndexTTS_F: IndexTTS/Export_IndexTTS.py:292-314,497-520; called at IndexTTS/Inference_IndexTTS_ONNX.py:787."""

    def __init__(self, engine, precision, state=None):
        from .config import INDEXTTS_VOCODER
        self.engine, self.precision = engine, precision
        self.cfg = INDEXTTS_VOCODER
        ch = self.cfg.stage_channels()
        self.inputs = tuple(NodeArg(f"save_bigvgan_conds_{i}", "tensor(float)", [1, c, 1]) for i, c in enumerate(ch)) + (
            NodeArg("bigvgan_cond_layer_speaker_embedding", "tensor(float)", [1, self.cfg.upsample_initial_channel, 1]),
            NodeArg("save_hidden_state", "tensor(float)", ["hidden_len", self.cfg.gpt_dim]))
        self.outputs = (NodeArg("generated_wav", "tensor(int16)", [1, 1, "generated_len"]),)
        state = state if state is not None else _checkpoints.get("indextts_f")
        if state is None:
            raise RuntimeError("no IndexTTS_F checkpoint registered: call session.register_checkpoint('indextts_f', state)")
        engine.load_state("ivgan", weights.ivgan_engine_tensors(state, self.cfg))
        engine.indextts_vocoder_build()

    def run(self, feed):
        conds = [_as_numpy(feed[f"save_bigvgan_conds_{i}"]).astype(np.float32) for i in range(len(self.cfg.upsample_rates))]
        cond_layer = _as_numpy(feed["bigvgan_cond_layer_speaker_embedding"]).astype(np.float32)
        hidden = _as_numpy(feed["save_hidden_state"]).astype(np.float32)
        if hidden.ndim != 2 or hidden.shape[1] != self.cfg.gpt_dim or hidden.shape[0] < 3:
            raise ValueError(f"save_hidden_state must be (S >= 3, {self.cfg.gpt_dim}), got {hidden.shape}")
        return [self.engine.indextts_vocoder_run(hidden, conds, cond_layer, precision=self.precision, hop=self.cfg.hop)]


_igpt_ready = {}


def load_indextts_gpt(engine, state=None, cfg=None):
    """Loads + builds the IndexTTS GPT-2 acoustic model once per engine (graphs B, C, E share it)."""
    from .config import INDEXTTS_GPT
    cfg = cfg or _checkpoints.get("indextts_gpt_config") or INDEXTTS_GPT
    key = id(engine)
    if state is None and key in _igpt_ready:
        return _igpt_ready[key]
    state = state if state is not None else _checkpoints.get("indextts_gpt")
    if state is None and _has_state(engine, "igpt"):
        pass                                 # uploaded from an engine blob (Engine.load_blob)
    elif state is None:
        raise RuntimeError("no IndexTTS GPT checkpoint registered: call session.register_checkpoint('indextts_gpt', state) or engine.load_blob(path)")
    else:
        engine.load_state("igpt", weights.igpt_engine_tensors(state, cfg))
    engine.indextts_gpt_build()
    _igpt_ready[key] = cfg
    return cfg


class _IndexTTSTextGraph(_Graph):
    """IndexTTS_B: IndexTTS/Export_IndexTTS.py:203-214,357-373; called at IndexTTS/Inference_IndexTTS_ONNX.py:723."""

    def __init__(self, engine, precision, state=None):
        self.engine = engine
        self.cfg = load_indextts_gpt(engine, state)
        self.inputs = (NodeArg("text_ids", "tensor(int32)", [1, "text_ids_len"]),)
        self.outputs = (NodeArg("text_hidden_state", "tensor(float)", [1, "text_ids_len", self.cfg.dim]),)

    def run(self, feed):
        return [self.engine.indextts_gpt_text_embed(_as_numpy(feed["text_ids"]).astype(np.int32))]


class _IndexTTSMelEmbedGraph(_Graph):
    """IndexTTS_C: Export_IndexTTS.py:217-225,377-390; called at Inference_IndexTTS_ONNX.py:729,775."""

    def __init__(self, engine, precision, state=None):
        self.engine = engine
        self.cfg = load_indextts_gpt(engine, state)
        self.inputs = (NodeArg("gpt_ids", "tensor(int32)", [1, 1]), NodeArg("kv_seq_len", "tensor(int64)", [1]))
        self.outputs = (NodeArg("gpt_hidden_state", "tensor(float)", [1, 1, self.cfg.dim]), NodeArg("next_kv_seq_len", "tensor(int64)", [1]))

    def run(self, feed):
        mel_id = int(_as_numpy(feed["gpt_ids"]).reshape(-1)[0])
        gen_len = int(_as_numpy(feed["kv_seq_len"]).reshape(-1)[0])
        h, nxt = self.engine.indextts_gpt_mel_embed(mel_id, gen_len)
        return [h, np.array([nxt], dtype=np.int64)]


class _IndexTTSConcatGraph(_Graph):
    """IndexTTS_D: Export_IndexTTS.py:228-235,394-414 -- a pure row concatenation (no arithmetic): host memory only."""

    def __init__(self, engine, precision, state=None):
        self.inputs = tuple(NodeArg(n, "tensor(float)", [1, f"{n}_len", "hidden"]) for n in ("embed_x", "embed_y", "embed_z"))
        self.outputs = (NodeArg("concat_hidden_state", "tensor(float)", [1, "concat_len", "hidden"]), NodeArg("concat_len", "tensor(int64)", [1]))

    def run(self, feed):
        h = np.concatenate([_as_numpy(feed[n]).astype(np.float32) for n in ("embed_x", "embed_y", "embed_z")], axis=1)
        return [h, np.array([h.shape[1]], dtype=np.int64)]


class ResidentKV(OrtValue):
    """out_key_<i> / out_value_<i> of IndexTTS_E: the cache stays in HBM; the reference loop only feeds these values back
    (Inference_IndexTTS_ONNX.py:766-767), so they are handles. ``numpy()`` materialises the reference layout on demand."""

    def __init__(self, engine, layer, which, rows, heads, owner=None, generation=0):
        self._engine, self._layer, self._which, self._rows, self._heads = engine, layer, which, rows, heads
        self._owner, self._generation = owner, generation      # the decode graph whose resident cache this handle views

    def is_current(self):
        """A handle is a VIEW of the resident cache: it is valid until the next IndexTTS_E call changes the cache."""
        return self._owner is None or self._owner._generation == self._generation

    @property
    def _array(self):                     # OrtValue.numpy(v) -- the reference calls it unbound -- reads this attribute
        if not self.is_current():
            raise RuntimeError("stale KV handle: the resident cache has advanced since this out_key / out_value was returned "
                               "(ORT would have kept a copy; read it with numpy() before the next IndexTTS_E call)")
        key, val = self._engine.indextts_gpt_kv_read(self._layer)
        return key if self._which == "key" else val

    def shape(self):
        return [self._heads, 64, self._rows] if self._which == "key" else [self._heads, self._rows, 64]


class _IndexTTSDecodeGraph(_Graph):
    """IndexTTS_E: Export_IndexTTS.py:238-289,416-482; called in the loop of Inference_IndexTTS_ONNX.py:753-781. The in_key / in_value
    feeds are not read (the cache is resident); history_len says whether the call continues it or starts a sentence."""

    def __init__(self, engine, precision, state=None):
        self.engine, self.precision = engine, precision
        self._generation = 0                 # bumped by every call: handles of older calls are stale views
        self.cfg = cfg = load_indextts_gpt(engine, state)
        L, H = cfg.layers, cfg.heads
        self.inputs = tuple(NodeArg(f"in_key_{i}", "tensor(float)", [H, 64, "history_len"]) for i in range(L)) + tuple(
            NodeArg(f"in_value_{i}", "tensor(float)", [H, "history_len", 64]) for i in range(L)) + (
            NodeArg("history_len", "tensor(int64)", [1]), NodeArg("repeat_penality", "tensor(float)", [1, cfg.mel_codes]),
            NodeArg("ids_len", "tensor(int64)", [1]), NodeArg("hidden_state", "tensor(float)", [1, "ids_len", cfg.dim]),
            NodeArg("attention_mask", "tensor(int8)", [1]))
        self.outputs = tuple(NodeArg(f"out_key_{i}", "tensor(float)", [H, 64, "kv_seq_len"]) for i in range(L)) + tuple(
            NodeArg(f"out_value_{i}", "tensor(float)", [H, "kv_seq_len", 64]) for i in range(L)) + (
            NodeArg("kv_seq_len", "tensor(int64)", [1]), NodeArg("last_hidden_state", "tensor(float)", [1, cfg.dim]),
            NodeArg("max_logit_id", "tensor(int32)", [1, 1]))

    def run(self, feed):
        cfg = self.cfg
        hidden = _as_numpy(feed["hidden_state"]).astype(np.float32)
        ids_len = int(_as_numpy(feed["ids_len"]).reshape(-1)[0])
        if hidden.ndim != 3 or hidden.shape[1] != ids_len or hidden.shape[2] != cfg.dim:
            raise ValueError(f"hidden_state must be (1, ids_len={ids_len}, {cfg.dim}), got {hidden.shape}")
        hist = int(_as_numpy(feed["history_len"]).reshape(-1)[0])
        flag = int(_as_numpy(feed["attention_mask"]).reshape(-1)[0])
        pen = _as_numpy(feed["repeat_penality"]).astype(np.float32)
        if hist > 0:
            # the in_key / in_value feeds are not uploaded (the cache is resident), so they must BE the resident cache: handles of
            # the latest call with history_len rows. ORT would honour any fed tensor; feeding older or edited caches is refused
            # here instead of silently attending over something else.
            for i in range(cfg.layers):
                for name in (f"in_key_{i}", f"in_value_{i}"):
                    v = feed.get(name)
                    if v is None:
                        continue
                    if not isinstance(v, ResidentKV):
                        raise ValueError(f"{name}: IndexTTS_E continues the RESIDENT cache; feed the out_key / out_value handles of the "
                                         f"previous call (got {type(v).__name__})")
                    if v._owner is not self or not v.is_current() or v._rows != hist:
                        raise ValueError(f"{name}: stale or foreign KV handle (rows {v._rows}, history_len {hist})")
        last, mid, kv = self.engine.indextts_gpt_step(hidden, hist, flag, pen, precision=self.precision)
        self._generation += 1
        keys = [ResidentKV(self.engine, i, "key", kv, cfg.heads, self, self._generation) for i in range(cfg.layers)]
        vals = [ResidentKV(self.engine, i, "value", kv, cfg.heads, self, self._generation) for i in range(cfg.layers)]
        return keys + vals + [np.array([kv], dtype=np.int64), last, mid]


_f5_ready = {}


def load_f5(engine, dit_state=None, vocos_state=None, cfg=F5):
    """Load + build the three F5 graphs' weights once per engine (the sessions share them)."""
    if _f5_ready.get(id(engine)):
        return
    dit_state = dit_state if dit_state is not None else _checkpoints.get("dit")
    vocos_state = vocos_state if vocos_state is not None else _checkpoints.get("vocos")
    if dit_state is None and vocos_state is None and _has_state(engine, "dit", "vocos", "f5"):
        engine.f5_build()                    # uploaded from an engine blob (Engine.load_blob)
        _f5_ready[id(engine)] = True
        return
    if dit_state is None or vocos_state is None:
        raise RuntimeError("no F5 checkpoint registered: call session.register_checkpoint('dit', ...) and ('vocos', ...)")
    engine.load_state("dit", weights.dit_engine_tensors(dit_state, cfg))
    engine.load_state("vocos", weights.vocos_engine_tensors(vocos_state, cfg))
    engine.load_state("f5", weights.f5_export_constants(dit_state, cfg))
    engine.f5_build()
    _f5_ready[id(engine)] = True


def _rope_rows(cfg, N):
    c = weights.f5_rope_rows(cfg)
    return c[0][:N], c[1][:N]


class _F5PreprocessGraph(_Graph):
    """F5_TTS/Export_F5.py:117-141, I/O names :294-306."""

    def __init__(self, engine, precision, state=None):
        self.engine, self.cfg = engine, F5
        load_f5(engine)
        f = "tensor(float)"
        self.inputs = (NodeArg("audio", "tensor(int16)", [1, 1, "audio_len"]), NodeArg("text_ids", "tensor(int32)", [1, "text_ids_len"]),
                       NodeArg("max_duration", "tensor(int64)", [1]))
        names = ["noise", "rope_cos_q", "rope_sin_q", "rope_cos_k", "rope_sin_k", "cat_mel_text", "cat_mel_text_drop", "ref_signal_len"]
        self.outputs = tuple(NodeArg(n, "tensor(int64)" if n == "ref_signal_len" else f, None) for n in names)

    def run(self, feed):
        cfg = self.cfg
        audio = _as_numpy(feed["audio"])
        text_ids = _as_numpy(feed["text_ids"])
        N = int(np.asarray(_as_numpy(feed["max_duration"])).reshape(-1)[0])
        if audio.dtype != np.int16 or audio.ndim != 3:
            raise ValueError("audio must be int16 (1, 1, L)")
        if not (0 < N <= cfg.max_frames):
            raise ValueError(f"max_duration must be in [1, {cfg.max_frames}]")
        cond, cond_drop, ref_len = self.engine.f5_preprocess(audio, text_ids, N, cfg.n_mels + cfg.text_dim)
        # noise = randn_like(zeros) (Export_F5.py:131): ORT's RandomNormalLike stream cannot be reproduced outside ORT;
        # a seeded numpy draw stands in (set_seed), and parity runs inject their own noise at graph B
        rng = np.random.default_rng(_seed[0])
        noise = rng.standard_normal((1, N, cfg.n_mels), dtype=np.float32)
        cos, sin = _rope_rows(cfg, N)
        rope_cos_q = np.broadcast_to(cos[None, None], (2, cfg.heads, N, cfg.head_dim))
        rope_sin_q = np.broadcast_to(sin[None, None], (2, cfg.heads, N, cfg.head_dim))
        return [noise, rope_cos_q, rope_sin_q, rope_cos_q.transpose(0, 1, 3, 2), rope_sin_q.transpose(0, 1, 3, 2),
                cond, cond_drop, np.array(ref_len, dtype=np.int64)]


class _F5TransformerGraph(_Graph):
    """F5_TTS/Export_F5.py:167-182 (FUSE_NFE = 1), I/O names :354-365. ``run_all_steps`` is the fused fast path;
    a loop of ``run`` gives the same numbers (same kernels, same order)."""

    def __init__(self, engine, precision, state=None):
        self.engine, self.cfg, self.precision = engine, F5, precision
        load_f5(engine)
        f = "tensor(float)"
        names = ["noise", "rope_cos_q", "rope_sin_q", "rope_cos_k", "rope_sin_k", "cat_mel_text", "cat_mel_text_drop"]
        self.inputs = tuple(NodeArg(n, f, None) for n in names) + (NodeArg("time_step", "tensor(int32)", [1]),)
        self.outputs = (NodeArg("denoised", f, [1, "max_duration", self.cfg.n_mels]), NodeArg("time_step", "tensor(int32)", [1]))

    def _step(self, feed, n_steps):
        noise = _as_numpy(feed["noise"]).astype(np.float32)
        cq, sq = _as_numpy(feed["rope_cos_q"]), _as_numpy(feed["rope_sin_q"])
        if cq.ndim != 4 or cq.shape[-1] != self.cfg.head_dim or cq.shape[2] != noise.shape[1]:
            raise ValueError("rope_cos_q must be (2, heads, max_duration, head_dim)")
        ts = int(np.asarray(_as_numpy(feed["time_step"])).reshape(-1)[0])
        out, ts2 = self.engine.f5_transformer(noise, np.ascontiguousarray(cq[0, 0], dtype=np.float32),
                                              np.ascontiguousarray(sq[0, 0], dtype=np.float32),
                                              _as_numpy(feed["cat_mel_text"]), _as_numpy(feed["cat_mel_text_drop"]), ts,
                                              n_steps=n_steps, precision=self.precision)
        return [out, np.array([ts2], dtype=np.int32)]

    def run(self, feed):
        return self._step(feed, 1)

    def run_all_steps(self, feed):
        ts = int(np.asarray(_as_numpy(feed["time_step"])).reshape(-1)[0])
        return self._step(feed, self.cfg.nfe - 1 - ts)


class _F5DecodeGraph(_Graph):
    """F5_TTS/Export_F5.py:193-203, I/O names :409-414."""

    def __init__(self, engine, precision, state=None):
        self.engine, self.cfg = engine, F5
        load_f5(engine)
        self.inputs = (NodeArg("denoised", "tensor(float)", [1, "max_duration", self.cfg.n_mels]),
                       NodeArg("ref_signal_len", "tensor(int64)", []))
        self.outputs = (NodeArg("output_audio", "tensor(int16)", [1, 1, "generated_len"]),)

    def run(self, feed):
        d = _as_numpy(feed["denoised"]).astype(np.float32)
        ref = int(np.asarray(_as_numpy(feed["ref_signal_len"])).reshape(-1)[0])
        if d.ndim != 3 or d.shape[2] != self.cfg.n_mels:
            raise ValueError("denoised must be (1, max_duration, n_mels)")
        return [self.engine.f5_decode(d, ref, hop=self.cfg.hop)]


_GRAPHS = {"bigvgan": _BigVGANGraph, "f5_preprocess": _F5PreprocessGraph, "f5_transformer": _F5TransformerGraph,
           "f5_decode": _F5DecodeGraph, "indextts_f": _IndexTTSVocoderGraph, "indextts_b": _IndexTTSTextGraph,
           "indextts_c": _IndexTTSMelEmbedGraph, "indextts_d": _IndexTTSConcatGraph, "indextts_e": _IndexTTSDecodeGraph}


def _kind_of(path: str) -> str:
    base = os.path.basename(str(path)).lower()
    for key in ("f5_preprocess", "f5_transformer", "f5_decode", "indextts_f", "indextts_b", "indextts_c", "indextts_d", "indextts_e",
                "bigvgan"):
        if key in base:
            return key
    if "indextts_a" in base:
        raise NotImplementedError(f"'{path}': the IndexTTS conditioning graph A (Export_IndexTTS.py:74-200) is not part of this engine "
                                  "(its conformer / perceiver / ECAPA modules are not in the reference tree, DESIGN.md 7): run it where the "
                                  "reference runs it, once per voice, and feed its outputs to the IndexTTS_D / IndexTTS_F sessions")
    raise ValueError(f"cannot tell which hot-path graph '{path}' is (expected BigVGAN / F5_Preprocess / F5_Transformer / F5_Decode / "
                     "IndexTTS_B .. IndexTTS_F in the file name)")


class InferenceSession:
    def __init__(self, path_or_bytes, sess_options=None, providers=None, provider_options=None, *,
                 device_id: int = 0, precision: str = "bf16", weights=None, **_ignored):
        self._kind = _kind_of(path_or_bytes)
        self._providers = ["B200ExecutionProvider"]
        # "fp16" is the arithmetic BASELINE.json's configs name (the reference's fp16 exports): its BigVGAN graph reports a float16
        # input, so a reference script feeds float16 mels (Export_BigVGAN.py:155-165), which run() widens again
        prec = {"fp32": capi.F32, "f32": capi.F32, "bf16": capi.BF16, "fp16": capi.F16, "f16": capi.F16}[precision]
        self._graph = _GRAPHS[self._kind](get_engine(device_id), prec, weights)
        self._inputs_meta = list(self._graph.inputs)
        self._outputs_meta = list(self._graph.outputs)

    def get_providers(self):
        return list(self._providers)

    def get_inputs(self):
        return list(self._inputs_meta)

    def get_outputs(self):
        return list(self._outputs_meta)

    def _select(self, output_names, outs):
        names = [o.name for o in self._outputs_meta]
        if not output_names:
            return outs
        return [outs[names.index(n)] for n in output_names]

    def run(self, output_names, input_feed, run_options=None):
        missing = [i.name for i in self._inputs_meta if i.name not in input_feed]
        if missing:
            raise ValueError(f"missing inputs: {missing}")
        return self._select(output_names, self._graph.run(input_feed))

    def run_with_ort_values(self, output_names, input_feed, run_options=None):
        return [o if isinstance(o, OrtValue) else OrtValue(o) for o in self.run(output_names, input_feed)]

    def io_binding(self):
        return IOBinding(self)

    def run_with_iobinding(self, iobinding, run_options=None):
        """Runs on the bound inputs and writes each bound output IN PLACE into its OrtValue (which may be one of the inputs)."""
        names = [o.name for o in self._outputs_meta]
        outs = self.run(names, dict(iobinding._inputs))
        for name, value in zip(names, outs):
            bound = iobinding._outputs.get(name)
            if bound is None:
                iobinding._outputs[name] = OrtValue(np.asarray(value))
            else:
                dst = bound.numpy()
                src = np.asarray(value)
                if dst.shape != src.shape or dst.dtype != src.dtype:
                    bound._array = src.copy()
                else:
                    dst[...] = src

    def run_all_steps(self, input_feed):
        """F5_Transformer only: all remaining NFE steps in one call, intermediates resident in HBM."""
        return self._graph.run_all_steps(input_feed)
