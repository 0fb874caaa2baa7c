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
from .config import BIGVGAN

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


def _as_numpy(v):
    return v.numpy() if isinstance(v, OrtValue) else np.asarray(v)


class _Graph:
    inputs = ()
    outputs = ()

    def run(self, feed: dict) -> list:
        raise NotImplementedError


class _BigVGANGraph(_Graph):
    """BigVGAN/Export_BigVGAN.py:37-49,65-70."""

    def __init__(self, engine, precision, state=None):
        self.engine, self.precision = engine, precision
        self.cfg = BIGVGAN
        self.inputs = (NodeArg("mel_features", "tensor(float)", [1, self.cfg.num_mels, "mel_features_len"]),)
        self.outputs = (NodeArg("generated_wav", "tensor(int16)", [1, 1, "generated_len"]),)
        state = state if state is not None else _checkpoints.get("bigvgan")
        if state is None:
            raise RuntimeError("no BigVGAN checkpoint registered: call session.register_checkpoint('bigvgan', state)")
        engine.load_state("bigvgan", weights.bigvgan_engine_tensors(state))
        engine.bigvgan_build()

    def run(self, feed):
        mel = _as_numpy(feed["mel_features"]).astype(np.float32)     # fp16 graphs feed float16
        if mel.ndim != 3 or mel.shape[1] != self.cfg.num_mels or mel.shape[2] == 0:
            raise ValueError(f"mel_features must be (B, {self.cfg.num_mels}, T>0), got {mel.shape}")
        return [self.engine.bigvgan_run(mel, precision=self.precision, hop=self.cfg.hop)]


_GRAPHS = {"bigvgan": _BigVGANGraph}


def _kind_of(path: str) -> str:
    base = os.path.basename(str(path)).lower()
    for key in ("f5_preprocess", "f5_transformer", "f5_decode", "bigvgan"):
        if key in base:
            return key
    raise ValueError(f"cannot tell which hot-path graph '{path}' is (expected BigVGAN / F5_Preprocess / "
                     "F5_Transformer / F5_Decode in the file name)")


class InferenceSession:
    def __init__(self, path_or_bytes, sess_options=None, providers=None, provider_options=None, *,
                 device_id: int = 0, precision: str = "bf16", weights=None, **_ignored):
        self._kind = _kind_of(path_or_bytes)
        self._providers = ["B200ExecutionProvider"]
        prec = {"fp32": capi.F32, "f32": capi.F32, "bf16": capi.BF16}[precision]
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
        return [OrtValue(o) for o in self.run(output_names, input_feed)]
