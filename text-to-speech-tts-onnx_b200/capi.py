"""ctypes binding of libb200tts.so (include/b200tts.h). No torch types cross this boundary.

The library is built in-tree by build.py; importing this module never builds anything and fails loudly
(`Libb200ttsMissing`) when the shared object is absent -- there is no CPU fallback.
"""
import ctypes
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200TTS_LIB") or os.path.join(HERE, "libb200tts.so")   # override: A/B runs of two builds

F32, BF16, F16 = 0, 1, 2          # include/b200tts.h: B200TTS_F32 / _BF16 / _F16


class Libb200ttsMissing(RuntimeError):
    pass


_c_f = ctypes.POINTER(ctypes.c_float)
_c_i16 = ctypes.POINTER(ctypes.c_int16)
_c_i32 = ctypes.POINTER(ctypes.c_int32)
_c_i64 = ctypes.POINTER(ctypes.c_int64)
_vp = ctypes.c_void_p
_int = ctypes.c_int

# name -> (restype, argtypes); every symbol include/b200tts.h declares (tests/test_capi_symbols.py checks the two agree)
SIGNATURES = {
    "b200tts_create": (_int, [_int, ctypes.POINTER(_vp)]),
    "b200tts_destroy": (None, [_vp]),
    "b200tts_last_error": (ctypes.c_char_p, []),
    "b200tts_set_stream": (_int, [_vp, _vp]),
    "b200tts_synchronize": (_int, [_vp]),
    "b200tts_set_option": (_int, [_vp, ctypes.c_char_p, _int]),
    "b200tts_launch_count": (ctypes.c_ulonglong, []),
    "b200tts_debug_chain_plan": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    "b200tts_load_tensor": (_int, [_vp, ctypes.c_char_p, _vp, _c_i64, _int]),
    "b200tts_load_tensor_device": (_int, [_vp, ctypes.c_char_p, _vp, _c_i64, _int]),
    "b200tts_bigvgan_build": (_int, [_vp]),
    "b200tts_bigvgan_run": (_int, [_vp, _vp, _int, _int, _int, _vp, _vp]),
    "b200tts_bigvgan_run_device": (_int, [_vp, _vp, _int, _int, _int, _vp, _vp]),
    "b200tts_indextts_vocoder_build": (_int, [_vp]),
    "b200tts_indextts_vocoder_run": (_int, [_vp, _vp, _int, _vp, _vp, _int, _vp, _vp, _vp]),
    "b200tts_indextts_gpt_build": (_int, [_vp]),
    "b200tts_indextts_gpt_info": (_int, [_vp, _c_i32, _c_i32, _c_i32, _c_i32, _c_i32]),
    "b200tts_indextts_gpt_text_embed": (_int, [_vp, _vp, _int, _vp]),
    "b200tts_indextts_gpt_mel_embed": (_int, [_vp, ctypes.c_int32, ctypes.c_int64, _vp]),
    "b200tts_indextts_gpt_step": (_int, [_vp, _vp, _int, ctypes.c_int64, _int, _vp, _int, _vp, _c_i32, _c_i64]),
    "b200tts_indextts_gpt_kv_read": (_int, [_vp, _int, _vp, _vp, _c_i64]),
    "b200tts_indextts_gpt_generate": (_int, [_vp, _vp, _int, _vp, _int, _int, _int, _vp, _vp, _vp, _c_i32]),
    "b200tts_indextts_gpt_generate_device": (_int, [_vp, _vp, _int, _vp, _int, _int, _int, _vp, _vp, _vp, _c_i32]),
    "b200tts_f5_build": (_int, [_vp]),
    "b200tts_f5_preprocess": (_int, [_vp, _vp, ctypes.c_int64, _vp, _int, ctypes.c_int64, _vp, _vp, _c_i64]),
    "b200tts_f5_transformer": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _int, _c_i32, _int, _int]),
    "b200tts_f5_decode": (_int, [_vp, _vp, _int, ctypes.c_int64, _vp, _vp, _c_i64]),
    "b200tts_f5_synthesize": (_int, [_vp, _vp, ctypes.c_int64, _vp, _int, ctypes.c_int64, _vp, _int, _int, _vp, _c_i64, _vp]),
    "b200tts_f5_synthesize_device": (_int, [_vp, _vp, ctypes.c_int64, _vp, _int, ctypes.c_int64, _vp, _int, _int, _vp, _vp]),
    "b200tts_f5_synthesize_batch_device": (_int, [_vp, _int, _vp, ctypes.c_int64, _vp, _int, ctypes.c_int64, _vp, _int, _int, _vp, _vp]),
    "b200tts_f5_bigvgan_pipeline": (_int, [_vp, _int, _vp, ctypes.c_int64, _vp, _int, ctypes.c_int64, _vp, _int, _int, _vp, _vp, _vp]),
    "b200tts_f5_bigvgan_pipeline_device": (_int, [_vp, _int, _vp, ctypes.c_int64, _vp, _int, ctypes.c_int64, _vp, _int, _int, _vp, _vp, _vp]),
    "b200tts_f5_bigvgan_pipeline_ragged": (_int, [_vp, _int, _vp, _c_i64, _vp, _c_i32, _c_i64, _vp, _int, _int, _vp, _vp, _vp]),
    "b200tts_f5_bigvgan_pipeline_ragged_device": (_int, [_vp, _int, _vp, _c_i64, _vp, _c_i32, _c_i64, _vp, _int, _int, _vp, _vp, _vp]),
    "b200tts_aa_activation": (_int, [_vp, _vp, _int, _int, _int, _vp, _vp, _vp, _int, _int, _vp]),
    "b200tts_conv1d": (_int, [_vp, _vp, _int, _int, _int, _vp, _int, _int, _int, _int, _vp, _int, _vp]),
    "b200tts_conv_transpose1d": (_int, [_vp, _vp, _int, _int, _int, _vp, _int, _int, _vp, _int, _vp]),
    "b200tts_attention": (_int, [_vp, _vp, _vp, _vp, _int, _int, _vp]),
    "b200tts_attention_prec": (_int, [_vp, _vp, _vp, _vp, _int, _int, _int, _vp]),
    "b200tts_bench_rowgemm": (_int, [_vp, _int, _int, _int, _int, _int, _int, _int, _int, _int, _c_f]),
    "b200tts_profile_begin": (_int, [_vp]),
    "b200tts_profile_end": (ctypes.c_char_p, [_vp]),
}

_lib = None


def load_library():
    """dlopen libb200tts.so and attach the prototypes. Works without a GPU (no CUDA call is made)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Libb200ttsMissing(
            f"{LIB_PATH} is missing: build it with `python text-to-speech-tts-onnx_b200/build.py` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Engine:
    """One engine per GPU (b200tts_create)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = _vp()
        rc = self.lib.b200tts_create(int(device), ctypes.byref(h))
        if rc != 0:
            raise RuntimeError("b200tts_create: " + self.lib.b200tts_last_error().decode())
        self.handle = h
        self.device = int(device)

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: " + self.lib.b200tts_last_error().decode())

    def close(self):
        if getattr(self, "handle", None):
            self.lib.b200tts_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing --------------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.b200tts_set_stream(self.handle, _vp(cuda_stream_ptr or 0)), "set_stream")

    def set_option(self, name: str, value: int):
        self._check(self.lib.b200tts_set_option(self.handle, name.encode(), int(value)), "set_option")

    def synchronize(self):
        self._check(self.lib.b200tts_synchronize(self.handle), "synchronize")

    def launch_count(self) -> int:
        return int(self.lib.b200tts_launch_count())

    def profile_begin(self):
        self._check(self.lib.b200tts_profile_begin(self.handle), "profile_begin")

    def profile_end(self) -> dict:
        return json.loads(self.lib.b200tts_profile_end(self.handle).decode())

    # -- weights ---------------------------------------------------------------------------------
    def load_tensor(self, name: str, array):
        a = _f32(array)
        shape = (ctypes.c_int64 * max(a.ndim, 1))(*a.shape)
        self._check(self.lib.b200tts_load_tensor(self.handle, name.encode(), _ptr(a), shape, a.ndim), f"load_tensor({name})")

    def load_tensor_device(self, name: str, dev_ptr: int, shape):
        shp = (ctypes.c_int64 * max(len(shape), 1))(*shape)
        self._check(self.lib.b200tts_load_tensor_device(self.handle, name.encode(), _vp(dev_ptr), shp, len(shape)),
                    f"load_tensor_device({name})")

    def load_state(self, prefix: str, state: dict):
        for k, v in state.items():
            self.load_tensor(f"{prefix}.{k}", v)
        self.__dict__.setdefault("loaded_prefixes", set()).add(prefix)      # session.py: a blob-loaded engine needs no re-upload

    def has_state(self, *prefixes) -> bool:
        """True when every given prefix ("dit", "vocos", "f5", "bigvgan", "igpt", "ivgan") has been uploaded (load_state / load_blob)."""
        have = self.__dict__.get("loaded_prefixes", set())
        return all(p in have for p in prefixes)

    def load_blob(self, path: str):
        """Upload every part of an engine blob written by checkpoint.save_blob / tools/convert_checkpoint.py."""
        from . import checkpoint
        return checkpoint.load_blob_into(self, path)

    def bigvgan_build(self):
        self._check(self.lib.b200tts_bigvgan_build(self.handle), "bigvgan_build")

    # -- BigVGAN ---------------------------------------------------------------------------------
    def bigvgan_run(self, mel, precision=F32, return_wave=False, hop=256, out=None):
        """out: optional caller-owned int16 array (B, 1, hop*T+30), e.g. pinned, to receive the PCM."""
        mel = _f32(mel)
        assert mel.ndim == 3, "mel_features must be (B, n_mels, T)"
        B, _, T = mel.shape
        n_out = hop * T + 30
        pcm = np.empty((B, 1, n_out), dtype=np.int16) if out is None else out
        assert pcm.dtype == np.int16 and pcm.size == B * n_out and pcm.flags["C_CONTIGUOUS"]
        wave = np.empty((B, 1, n_out), dtype=np.float32) if return_wave else None
        self._check(self.lib.b200tts_bigvgan_run(self.handle, _ptr(mel), B, T, int(precision), _ptr(pcm), _ptr(wave)),
                    "bigvgan_run")
        return (pcm, wave) if return_wave else pcm

    def bigvgan_run_device(self, mel_ptr: int, B: int, T: int, pcm_ptr: int, precision=BF16, wave_ptr: int = 0):
        self._check(self.lib.b200tts_bigvgan_run_device(self.handle, _vp(mel_ptr), B, T, int(precision), _vp(pcm_ptr),
                                                        _vp(wave_ptr or 0)), "bigvgan_run_device")

    # -- IndexTTS_F vocoder ----------------------------------------------------------------------
    def indextts_vocoder_build(self):
        self._check(self.lib.b200tts_indextts_vocoder_build(self.handle), "indextts_vocoder_build")

    def indextts_vocoder_run(self, hidden, conds, cond_layer, precision=F32, return_wave=False, hop=1024):
        """save_hidden_state (S, gpt_dim), [save_bigvgan_conds_i], bigvgan_cond_layer_speaker_embedding ->
        generated_wav int16 (1, 1, hop*(S-2)+30)."""
        hidden = _f32(hidden)
        assert hidden.ndim == 2 and hidden.shape[0] >= 3, "save_hidden_state must be (S >= 3, gpt_dim)"
        S = hidden.shape[0]
        cs = [_f32(np.asarray(c).reshape(-1)) for c in conds]
        cl = _f32(np.asarray(cond_layer).reshape(-1))
        ptrs = (ctypes.c_void_p * len(cs))(*[c.ctypes.data for c in cs])
        n_out = hop * (S - 2) + 30
        pcm = np.empty((1, 1, n_out), dtype=np.int16)
        wave = np.empty((1, 1, n_out), dtype=np.float32) if return_wave else None
        got = ctypes.c_int64(0)
        self._check(self.lib.b200tts_indextts_vocoder_run(self.handle, _ptr(hidden), S, ptrs, _ptr(cl), int(precision), _ptr(pcm),
                                                          _ptr(wave), ctypes.byref(got)), "indextts_vocoder_run")
        assert got.value == n_out, (got.value, n_out)
        return (pcm, wave) if return_wave else pcm

    # -- IndexTTS GPT-2 acoustic model (graphs B-E) -----------------------------------------------
    def indextts_gpt_build(self):
        self._check(self.lib.b200tts_indextts_gpt_build(self.handle), "indextts_gpt_build")

    def indextts_gpt_info(self) -> dict:
        v = [ctypes.c_int32(0) for _ in range(5)]
        self._check(self.lib.b200tts_indextts_gpt_info(self.handle, *[ctypes.byref(x) for x in v]), "indextts_gpt_info")
        return dict(zip(("dim", "layers", "heads", "mel_codes", "max_rows"), (x.value for x in v)))

    def indextts_gpt_text_embed(self, text_ids):
        """graph B: text_ids (1, n) int32 -> text_hidden_state (1, n + 2, dim)."""
        ids = np.ascontiguousarray(np.asarray(text_ids, dtype=np.int32).reshape(-1))
        D = self.indextts_gpt_info()["dim"]
        out = np.empty((1, ids.size + 2, D), dtype=np.float32)
        self._check(self.lib.b200tts_indextts_gpt_text_embed(self.handle, _ptr(ids), int(ids.size), _ptr(out)), "indextts_gpt_text_embed")
        return out

    def indextts_gpt_mel_embed(self, mel_id: int, gen_len: int):
        """graph C: -> (gpt_hidden_state (1, 1, dim), gen_len + 1)."""
        D = self.indextts_gpt_info()["dim"]
        out = np.empty((1, 1, D), dtype=np.float32)
        self._check(self.lib.b200tts_indextts_gpt_mel_embed(self.handle, int(mel_id), int(gen_len), _ptr(out)), "indextts_gpt_mel_embed")
        return out, int(gen_len) + 1

    def indextts_gpt_step(self, hidden, history_len: int, attention_mask: int, repeat_penality, precision=F32):
        """graph E on the resident KV cache -> (last_hidden_state (1, dim), max_logit_id (1, 1) int32, kv_seq_len)."""
        hidden = _f32(hidden)
        D = hidden.shape[-1]
        rows = hidden.size // D
        pen = _f32(np.asarray(repeat_penality).reshape(-1))
        last = np.empty((1, D), dtype=np.float32)
        mid = ctypes.c_int32(0)
        kv = ctypes.c_int64(0)
        self._check(self.lib.b200tts_indextts_gpt_step(self.handle, _ptr(hidden), int(rows), int(history_len), int(attention_mask),
                                                       _ptr(pen), int(precision), _ptr(last), ctypes.byref(mid), ctypes.byref(kv)),
                    "indextts_gpt_step")
        return last, np.array([[mid.value]], dtype=np.int32), int(kv.value)

    def indextts_gpt_kv_read(self, layer: int):
        """-> (out_key_<layer> (H, 64, S), out_value_<layer> (H, S, 64)) of the resident cache."""
        info = self.indextts_gpt_info()
        rows = ctypes.c_int64(0)
        self._check(self.lib.b200tts_indextts_gpt_kv_read(self.handle, int(layer), None, None, ctypes.byref(rows)), "indextts_gpt_kv_read")
        S, H = int(rows.value), info["heads"]
        key = np.zeros((H, 64, S), dtype=np.float32)
        val = np.zeros((H, S, 64), dtype=np.float32)
        if S:
            self._check(self.lib.b200tts_indextts_gpt_kv_read(self.handle, int(layer), _ptr(key), _ptr(val), ctypes.byref(rows)),
                        "indextts_gpt_kv_read")
        return key, val

    def indextts_gpt_generate(self, conds_latent, text_ids, max_new: int = 0, precision=BF16, penalty=None):
        """One sentence, loop on the device -> (ids (n,) int32, hidden (n, dim) f32[, penalty (1, mel_codes) when one was passed])."""
        info = self.indextts_gpt_info()
        D = info["dim"]
        conds = _f32(np.asarray(conds_latent).reshape(-1, D))
        ids = np.ascontiguousarray(np.asarray(text_ids, dtype=np.int32).reshape(-1))
        cap = info["max_rows"] + 1
        out_ids = np.zeros((cap,), dtype=np.int32)
        out_hid = np.zeros((cap, D), dtype=np.float32)
        pen = None if penalty is None else np.ascontiguousarray(np.asarray(penalty, dtype=np.float32).reshape(1, -1)).copy()
        n = ctypes.c_int32(0)
        self._check(self.lib.b200tts_indextts_gpt_generate(self.handle, _ptr(conds), int(conds.shape[0]), _ptr(ids), int(ids.size),
                                                           int(max_new), int(precision), _ptr(pen), _ptr(out_ids), _ptr(out_hid),
                                                           ctypes.byref(n)), "indextts_gpt_generate")
        res = (out_ids[:n.value].copy(), out_hid[:n.value].copy())
        return res + (pen,) if pen is not None else res

    def indextts_gpt_generate_device(self, conds_ptr, cond_rows, ids_ptr, n_text, ids_out_ptr, hidden_out_ptr, max_new=0, precision=BF16,
                                     penalty_ptr=0) -> int:
        n = ctypes.c_int32(0)
        self._check(self.lib.b200tts_indextts_gpt_generate_device(self.handle, _vp(conds_ptr), int(cond_rows), _vp(ids_ptr), int(n_text),
                                                                  int(max_new), int(precision), _vp(penalty_ptr) if penalty_ptr else None,
                                                                  _vp(ids_out_ptr), _vp(hidden_out_ptr), ctypes.byref(n)),
                    "indextts_gpt_generate_device")
        return int(n.value)

    # -- F5-TTS ----------------------------------------------------------------------------------
    def f5_build(self):
        self._check(self.lib.b200tts_f5_build(self.handle), "f5_build")

    def f5_preprocess(self, audio, text_ids, max_duration: int, cond_dim: int = 612):
        """audio int16 (1,1,L), text_ids int32 (1,n) -> cat_mel_text, cat_mel_text_drop (1,N,cond_dim), ref_signal_len."""
        audio = np.ascontiguousarray(audio, dtype=np.int16).reshape(-1)
        ids = np.ascontiguousarray(text_ids, dtype=np.int32).reshape(-1)
        N = int(max_duration)
        c = np.empty((1, N, cond_dim), dtype=np.float32)
        cd = np.empty((1, N, cond_dim), dtype=np.float32)
        ref = ctypes.c_int64(0)
        self._check(self.lib.b200tts_f5_preprocess(self.handle, _ptr(audio), audio.size, _ptr(ids), ids.size, N, _ptr(c),
                                                   _ptr(cd), ctypes.byref(ref)), "f5_preprocess")
        return c, cd, int(ref.value)

    def f5_transformer(self, noise, rope_cos, rope_sin, cond, cond_drop, time_step: int, n_steps: int = 1, precision=F32):
        """One (or n_steps fused) NFE step(s). noise (1,N,100) is copied, updated and returned with the new time_step."""
        noise = np.array(noise, dtype=np.float32, order="C", copy=True)
        N = noise.shape[-2]
        rc, rs, c, cd = _f32(rope_cos), _f32(rope_sin), _f32(cond), _f32(cond_drop)
        assert rc.shape[-2:] == (N, 64) and rc.size == N * 64 and c.shape[-2] == N and cd.shape == c.shape
        ts = ctypes.c_int32(int(time_step))
        self._check(self.lib.b200tts_f5_transformer(self.handle, _ptr(noise), _ptr(rc), _ptr(rs), _ptr(c), _ptr(cd), N,
                                                    ctypes.byref(ts), int(n_steps), int(precision)), "f5_transformer")
        return noise, int(ts.value)

    def f5_decode(self, denoised, ref_signal_len: int, return_wave=False, hop: int = 256):
        d = _f32(denoised)
        N = d.shape[-2]
        ns = max(hop * (N - int(ref_signal_len) - 1), 0)
        pcm = np.empty((1, 1, max(ns, 1)), dtype=np.int16)
        wave = np.empty((1, 1, max(ns, 1)), dtype=np.float32) if return_wave else None
        n_out = ctypes.c_int64(0)
        self._check(self.lib.b200tts_f5_decode(self.handle, _ptr(d), N, int(ref_signal_len), _ptr(pcm), _ptr(wave),
                                               ctypes.byref(n_out)), "f5_decode")
        assert n_out.value == ns
        pcm = pcm[..., :ns]
        return (pcm, wave[..., :ns]) if return_wave else pcm

    def f5_synthesize(self, audio, text_ids, max_duration: int, noise, precision=BF16, n_steps: int = -1, return_mel=False,
                      hop: int = 256):
        audio = np.ascontiguousarray(audio, dtype=np.int16).reshape(-1)
        ids = np.ascontiguousarray(text_ids, dtype=np.int32).reshape(-1)
        noise = _f32(noise)
        N = int(max_duration)
        assert noise.size == N * 100
        ns = hop * (N - (audio.size // hop + 1) - 1)
        pcm = np.empty((1, 1, max(ns, 0)), dtype=np.int16)
        mel = np.empty((1, N, 100), dtype=np.float32) if return_mel else None
        n_out = ctypes.c_int64(0)
        self._check(self.lib.b200tts_f5_synthesize(self.handle, _ptr(audio), audio.size, _ptr(ids), ids.size, N, _ptr(noise),
                                                   int(precision), int(n_steps), _ptr(pcm), ctypes.byref(n_out), _ptr(mel)),
                    "f5_synthesize")
        return (pcm, mel) if return_mel else pcm

    def f5_synthesize_device(self, audio_ptr, L, ids_ptr, n_text, max_duration, noise_ptr, pcm_ptr, precision=BF16,
                             n_steps: int = -1, mel_ptr: int = 0):
        self._check(self.lib.b200tts_f5_synthesize_device(self.handle, _vp(audio_ptr), int(L), _vp(ids_ptr), int(n_text),
                                                          int(max_duration), _vp(noise_ptr), int(precision), int(n_steps),
                                                          _vp(pcm_ptr), _vp(mel_ptr or 0)), "f5_synthesize_device")

    def f5_synthesize_batch_device(self, U, audio_ptr, L, ids_ptr, n_text, max_duration, noise_ptr, pcm_ptr, precision=BF16,
                                   n_steps: int = -1, mel_ptr: int = 0):
        """U utterances sharing (L, n_text, max_duration) through one batched DiT loop (device buffers, no sync)."""
        self._check(self.lib.b200tts_f5_synthesize_batch_device(self.handle, int(U), _vp(audio_ptr), int(L), _vp(ids_ptr), int(n_text),
                                                                int(max_duration), _vp(noise_ptr), int(precision), int(n_steps),
                                                                _vp(pcm_ptr), _vp(mel_ptr or 0)), "f5_synthesize_batch_device")

    def f5_bigvgan_pipeline(self, audio, text_ids, max_duration: int, noise, precision=BF16, n_steps: int = -1, with_vocos=False,
                            return_mel=False, hop: int = 256, out=None):
        """U utterances sharing (L, n_text, max_duration): audio (U, L) i16, text_ids (U, n_text) i32, noise (U, N, 100) f32
        -> BigVGAN wav (U, 256*G+30) i16 [, Vocos wav (U, 256*(G-1))] [, mel (U, N, 100)] through ONE host-buffer C call."""
        audio = np.ascontiguousarray(audio, dtype=np.int16)
        audio = audio.reshape(-1, audio.shape[-1])
        U, L = audio.shape
        ids = np.ascontiguousarray(text_ids, dtype=np.int32).reshape(U, -1)
        N = int(max_duration)
        noise = _f32(noise).reshape(U, N, 100)
        G = N - (L // hop + 1)
        wav = np.empty((U, hop * G + 30), dtype=np.int16) if out is None else out      # out: caller-owned (e.g. pinned) PCM buffer
        assert wav.dtype == np.int16 and wav.size == U * (hop * G + 30) and wav.flags["C_CONTIGUOUS"]
        voc = np.empty((U, hop * (G - 1)), dtype=np.int16) if with_vocos else None
        mel = np.empty((U, N, 100), dtype=np.float32) if return_mel else None
        self._check(self.lib.b200tts_f5_bigvgan_pipeline(self.handle, U, _ptr(audio), L, _ptr(ids), ids.shape[1], N, _ptr(noise),
                                                         int(precision), int(n_steps), _ptr(wav), _ptr(voc), _ptr(mel)),
                    "f5_bigvgan_pipeline")
        res = [wav]
        if with_vocos:
            res.append(voc)
        if return_mel:
            res.append(mel)
        return res[0] if len(res) == 1 else tuple(res)

    def f5_bigvgan_pipeline_device(self, U, audio_ptr, L, ids_ptr, n_text, max_duration, noise_ptr, wav_ptr, precision=BF16,
                                   n_steps: int = -1, wav_vocos_ptr: int = 0, mel_ptr: int = 0):
        self._check(self.lib.b200tts_f5_bigvgan_pipeline_device(self.handle, int(U), _vp(audio_ptr), int(L), _vp(ids_ptr), int(n_text),
                                                                int(max_duration), _vp(noise_ptr), int(precision), int(n_steps),
                                                                _vp(wav_ptr), _vp(wav_vocos_ptr or 0), _vp(mel_ptr or 0)),
                    "f5_bigvgan_pipeline_device")

    def f5_bigvgan_pipeline_ragged(self, audios, text_ids, max_durations, noises, precision=BF16, n_steps: int = -1, with_vocos=False,
                                   return_mel=False, hop: int = 256, out=None):
        """Ragged batch: lists of per-utterance arrays -- audios[u] (L_u,) i16, text_ids[u] (n_u,) i32, max_durations[u], noises[u]
        (N_u, 100) f32 -> list of BigVGAN wavs (256*G_u + 30,) [, list of Vocos wavs] [, list of mels], ONE host-buffer C call."""
        U = len(audios)
        au = [np.ascontiguousarray(a, dtype=np.int16).reshape(-1) for a in audios]
        tx = [np.ascontiguousarray(t, dtype=np.int32).reshape(-1) for t in text_ids]
        Ns = np.asarray([int(n) for n in max_durations], dtype=np.int64)
        nz = [_f32(z).reshape(int(n), 100) for z, n in zip(noises, Ns)]
        L = np.asarray([a.size for a in au], dtype=np.int64)
        nt = np.asarray([t.size for t in tx], dtype=np.int32)
        G = Ns - (L // hop + 1)
        nv, ns = hop * G + 30, hop * (G - 1)
        a_cat, t_cat, z_cat = np.concatenate(au), np.concatenate(tx), np.concatenate(nz, 0)
        wav = np.empty(int(nv.sum()), dtype=np.int16) if out is None else out
        voc = np.empty(int(ns.sum()), dtype=np.int16) if with_vocos else None
        mel = np.empty((int(Ns.sum()), 100), dtype=np.float32) if return_mel else None
        self.f5_bigvgan_pipeline_ragged_concat(a_cat, L, t_cat, nt, Ns, z_cat, wav, precision=precision, n_steps=n_steps, voc=voc, mel=mel)
        split = lambda x, sizes: [x[int(o):int(o + n)] for o, n in zip(np.concatenate([[0], np.cumsum(sizes)[:-1]]), sizes)]
        res = [split(wav, nv)]
        if with_vocos:
            res.append(split(voc, ns))
        if return_mel:
            res.append(split(mel, Ns))
        return res[0] if len(res) == 1 else tuple(res)

    def f5_bigvgan_pipeline_ragged_concat(self, audio_cat, L, ids_cat, n_text, max_duration, noise_cat, wav, precision=BF16, n_steps: int = -1,
                                          voc=None, mel=None):
        """The C call itself on already concatenated host buffers (bench.py: pinned): audio_cat i16 [sum L], ids_cat i32 [sum n_text],
        noise_cat f32 [sum N, 100]; L / max_duration int64 [U], n_text int32 [U]; wav i16 [sum (256 G + 30)] (caller-owned)."""
        L = np.ascontiguousarray(L, dtype=np.int64)
        n_text = np.ascontiguousarray(n_text, dtype=np.int32)
        max_duration = np.ascontiguousarray(max_duration, dtype=np.int64)
        assert audio_cat.dtype == np.int16 and ids_cat.dtype == np.int32 and noise_cat.dtype == np.float32 and wav.dtype == np.int16
        assert audio_cat.size == int(L.sum()) and ids_cat.size == int(n_text.sum()) and noise_cat.size == int(max_duration.sum()) * 100
        self._check(self.lib.b200tts_f5_bigvgan_pipeline_ragged(self.handle, int(L.size), _ptr(audio_cat), L.ctypes.data_as(_c_i64), _ptr(ids_cat),
                                                                n_text.ctypes.data_as(_c_i32), max_duration.ctypes.data_as(_c_i64),
                                                                _ptr(noise_cat), int(precision), int(n_steps), _ptr(wav), _ptr(voc), _ptr(mel)),
                    "f5_bigvgan_pipeline_ragged")

    def f5_bigvgan_pipeline_ragged_device(self, U, audio_ptr, L, ids_ptr, n_text, max_duration, noise_ptr, wav_ptr, precision=BF16,
                                          n_steps: int = -1, wav_vocos_ptr: int = 0, mel_ptr: int = 0):
        """L, n_text, max_duration: host numpy arrays (int64, int32, int64) of U entries; the data pointers are device pointers."""
        L = np.ascontiguousarray(L, dtype=np.int64)
        n_text = np.ascontiguousarray(n_text, dtype=np.int32)
        max_duration = np.ascontiguousarray(max_duration, dtype=np.int64)
        self._check(self.lib.b200tts_f5_bigvgan_pipeline_ragged_device(self.handle, int(U), _vp(audio_ptr), L.ctypes.data_as(_c_i64),
                                                                       _vp(ids_ptr), n_text.ctypes.data_as(_c_i32),
                                                                       max_duration.ctypes.data_as(_c_i64), _vp(noise_ptr), int(precision),
                                                                       int(n_steps), _vp(wav_ptr), _vp(wav_vocos_ptr or 0), _vp(mel_ptr or 0)),
                    "f5_bigvgan_pipeline_ragged_device")

    def bench_rowgemm(self, B, M, N, Cin, taps=1, dil=1, groups=1, epilogue=0, iters=20) -> float:
        """Average ms per launch of the tensor-core shifted-row GEMM on synthetic operands (tools/bench_gemm.py)."""
        ms = ctypes.c_float(0)
        self._check(self.lib.b200tts_bench_rowgemm(self.handle, B, M, N, Cin, taps, dil, groups, epilogue, iters, ctypes.byref(ms)),
                    "bench_rowgemm")
        return float(ms.value)

    # -- single ops (reference layouts) --------------------------------------------------------------
    def aa_activation(self, x, alpha_log, beta_log, taps12, precise=True, post=False):
        x = _f32(x)
        B, C, L = x.shape
        y = np.empty((B, C, L + 30 if post else L), dtype=np.float32)
        a, b, t = _f32(alpha_log), _f32(beta_log), _f32(taps12)
        assert a.shape == (C,) and b.shape == (C,) and t.shape == (12,)
        self._check(self.lib.b200tts_aa_activation(self.handle, _ptr(x), B, C, L, _ptr(a), _ptr(b), _ptr(t),
                                                   int(precise), int(post), _ptr(y)), "aa_activation")
        return y

    def attention(self, q, k, v, precision=BF16):
        """q, k, v (2, H, N, 64) fp32 -> softmax(q k^T) v as (2, N, H*64), tcgen05 path (bf16 or fp16 operands)."""
        q, k, v = _f32(q), _f32(k), _f32(v)
        two, H, N, hd = q.shape
        assert two == 2 and hd == 64 and k.shape == q.shape and v.shape == q.shape
        out = np.empty((2, N, H * 64), dtype=np.float32)
        self._check(self.lib.b200tts_attention_prec(self.handle, _ptr(q), _ptr(k), _ptr(v), H, N, int(precision), _ptr(out)),
                    "attention")
        return out

    def conv1d(self, x, w, bias=None, dilation=1, groups=1, precision=F32):
        x, w = _f32(x), _f32(w)
        B, Cin, L = x.shape
        Cout, cg, k = w.shape
        assert cg * groups == Cin
        b = None if bias is None else _f32(bias)
        y = np.empty((B, Cout, L), dtype=np.float32)
        self._check(self.lib.b200tts_conv1d(self.handle, _ptr(x), B, Cin, L, _ptr(w), Cout, k, int(dilation), int(groups),
                                            _ptr(b), int(precision), _ptr(y)), "conv1d")
        return y

    def conv_transpose1d(self, x, w, bias=None, stride=2, precision=F32):
        x, w = _f32(x), _f32(w)
        B, Cin, L = x.shape
        cin2, Cout, k = w.shape
        assert cin2 == Cin and k == 2 * stride
        b = None if bias is None else _f32(bias)
        y = np.empty((B, Cout, L * stride), dtype=np.float32)
        self._check(self.lib.b200tts_conv_transpose1d(self.handle, _ptr(x), B, Cin, L, _ptr(w), Cout, int(stride), _ptr(b),
                                                      int(precision), _ptr(y)), "conv_transpose1d")
        return y
