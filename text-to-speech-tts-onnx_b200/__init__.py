"""b200tts: B200-native (sm_100a) engine for the F5-TTS / BigVGAN hot path of
DakeQQ/Text-to-Speech-TTS-ONNX, behind the reference's InferenceSession call surface.

Import as ``import b200tts`` (see b200tts.py at the repo root)."""
from . import config, synth  # noqa: F401

__all__ = ["config", "synth"]
