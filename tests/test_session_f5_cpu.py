"""CPU: host logic of the onnxruntime-shaped F5 sessions (session.py: F5_Preprocess / F5_Transformer / F5_Decode, the IOBinding
loop) driven as F5-TTS-ONNX-Inference.py:247-311 drives them, with the device engine replaced by a stand-in that answers the
capi calls from the oracle. Two Euler steps on a short utterance keep it to seconds; the arithmetic is the GPU tests' job."""
import numpy as np
import pytest
import torch

import b200tts  # noqa: F401
from b200tts import config, session as ort, synth
from oracle import f5_ref as R

CFG = config.F5
STEPS = 2


class FakeEngine:
    def __init__(self, dit_sd, vocos_sd):
        self.sd = R.prescale_qk(dit_sd, CFG)
        self.tables = R.time_tables(self.sd, CFG)
        self.fsd = R.fold_vocos(vocos_sd, CFG)
        self.loaded, self.steps_run = [], 0

    def load_state(self, prefix, tensors):
        self.loaded.append(prefix)

    def f5_build(self):
        pass

    @torch.inference_mode()
    def f5_preprocess(self, audio, text_ids, N, cond_dim):
        out = R.f5_preprocess(audio, text_ids, [N], self.sd, CFG, noise=np.zeros((1, N, CFG.n_mels), np.float32))
        assert out[5].shape[-1] == cond_dim
        return out[5].numpy(), out[6].numpy(), out[7]

    @torch.inference_mode()
    def f5_transformer(self, noise, cos, sin, cond, cond_drop, ts, n_steps=1, precision=0):
        x = torch.from_numpy(noise)
        for _ in range(n_steps):
            x, ts = R.f5_transformer_step(self.sd, x, torch.from_numpy(np.asarray(cond)), torch.from_numpy(np.asarray(cond_drop)), ts,
                                          self.tables, CFG, torch.from_numpy(np.array(cos)), torch.from_numpy(np.array(sin)))
            self.steps_run += 1
        return x.numpy(), ts

    @torch.inference_mode()
    def f5_decode(self, d, ref, hop=256):
        return R.f5_decode(d, ref, self.fsd, CFG).numpy()


@pytest.fixture(scope="module")
def setup():
    dit_sd, vocos_sd = synth.f5_dit_state(4321), synth.vocos_state(2468)
    audio, text_ids, maxd, noise = synth.f5_inputs(5, 12800, 12)           # 51 reference frames, N ~ 100
    want = R.f5_synthesize(audio, text_ids, maxd, noise, dit_sd, vocos_sd, CFG, steps=STEPS).numpy()
    return dit_sd, vocos_sd, audio, text_ids, maxd, noise, want


@pytest.fixture()
def sessions(setup, monkeypatch):
    fake = FakeEngine(setup[0], setup[1])
    monkeypatch.setattr(ort, "get_engine", lambda device=0: fake)
    ort.register_checkpoint("dit", setup[0])
    ort.register_checkpoint("vocos", setup[1])
    ort._f5_ready.clear()
    s = [ort.InferenceSession(f"/m/F5_{k}.onnx", precision="fp32") for k in ("Preprocess", "Transformer", "Decode")]
    yield fake, s
    ort._f5_ready.clear()
    for k in ("dit", "vocos"):
        ort._checkpoints.pop(k, None)


def test_io_names_match_the_export(sessions):
    fake, (A, B, C) = sessions
    assert fake.loaded == ["dit", "vocos", "f5"]                     # one load for the three graphs
    assert [i.name for i in A.get_inputs()] == ["audio", "text_ids", "max_duration"]                                  # Export_F5.py:294-306
    assert [o.name for o in A.get_outputs()] == ["noise", "rope_cos_q", "rope_sin_q", "rope_cos_k", "rope_sin_k", "cat_mel_text",
                                                 "cat_mel_text_drop", "ref_signal_len"]
    assert [i.name for i in B.get_inputs()] == [o.name for o in A.get_outputs()][:7] + ["time_step"]                   # :354-365
    assert [o.name for o in B.get_outputs()] == ["denoised", "time_step"]
    assert [i.name for i in C.get_inputs()] == ["denoised", "ref_signal_len"] and [o.name for o in C.get_outputs()] == ["output_audio"]
    assert "float16" not in B._inputs_meta[0].type and A.get_providers() == ["B200ExecutionProvider"]


def test_reference_host_loop(sessions, setup):
    fake, (A, B, C) = sessions
    _, _, audio, text_ids, maxd, noise0, want = setup
    outA = [o.name for o in A.get_outputs()]
    noise, cq, sq, ck, sk, cat, cat_drop, ref_len = A.run(outA, {"audio": audio, "text_ids": text_ids, "max_duration": maxd})
    N = int(maxd[0])
    assert noise.shape == (1, N, CFG.n_mels) and cq.shape == (2, CFG.heads, N, CFG.head_dim) and ck.shape == (2, CFG.heads, CFG.head_dim, N)
    assert cat.shape == (1, N, CFG.n_mels + CFG.text_dim) and ref_len.dtype == np.int64 and int(ref_len) == 12800 // CFG.hop + 1
    noise = noise0                                                      # parity runs inject the Euler start (session.py note)
    time_step = np.array([0], dtype=np.int32)
    inB = [i.name for i in B.get_inputs()]
    for _ in range(STEPS):
        noise, time_step = B.run(["denoised", "time_step"], dict(zip(inB, [noise, cq, sq, ck, sk, cat, cat_drop, time_step])))
        assert time_step.dtype == np.int32 and time_step.shape == (1,)
    assert int(time_step[0]) == STEPS and fake.steps_run == STEPS
    pcm = C.run(["output_audio"], {"denoised": noise, "ref_signal_len": ref_len})[0]
    assert pcm.dtype == np.int16 and pcm.shape == want.shape
    np.testing.assert_array_equal(pcm, want)


def test_iobinding_loop_updates_in_place(sessions, setup):
    """The dormant bound loop (F5-TTS-ONNX-Inference.py:257-288): outputs aliased onto the inputs `noise` and `time_step`."""
    fake, (A, B, C) = sessions
    _, _, audio, text_ids, maxd, noise0, want = setup
    outs = A.run(None, {"audio": audio, "text_ids": text_ids, "max_duration": maxd})
    OV = ort.OrtValue
    vals = [OV.ortvalue_from_numpy(noise0.copy())] + [OV.ortvalue_from_numpy(np.ascontiguousarray(o)) for o in outs[1:7]] + [
        OV.ortvalue_from_numpy(np.array([0], dtype=np.int32))]
    io = B.io_binding()
    for node, v in zip(B.get_inputs(), vals):
        io.bind_ortvalue_input(name=node.name, ortvalue=v)
    for node, v in zip(B.get_outputs(), [vals[0], vals[-1]]):
        io.bind_ortvalue_output(name=node.name, ortvalue=v)
    for _ in range(STEPS):
        B.run_with_iobinding(io)
    assert int(OV.numpy(vals[-1])[0]) == STEPS                          # time_step advanced through the alias
    noise = OV.numpy(io.get_outputs()[0])
    assert noise is OV.numpy(vals[0])
    pcm = C.run(None, {"denoised": noise, "ref_signal_len": outs[7]})[0]
    np.testing.assert_array_equal(pcm, want)


def test_run_all_steps_counts_the_remaining_steps(sessions, setup):
    fake, (A, B, C) = sessions
    _, _, audio, text_ids, maxd, noise0, _ = setup
    outs = A.run(None, {"audio": audio, "text_ids": text_ids, "max_duration": maxd})
    feed = dict(zip([i.name for i in B.get_inputs()], [noise0] + list(outs[1:7]) + [np.array([CFG.nfe - 3], np.int32)]))
    _, ts = B.run_all_steps(feed)
    assert fake.steps_run == 2 and int(ts[0]) == CFG.nfe - 1            # NFE - 1 Euler steps in total (quirk q10)
    with pytest.raises(ValueError, match="max_duration"):
        A.run(None, {"audio": audio, "text_ids": text_ids, "max_duration": np.array([CFG.max_frames + 1])})
