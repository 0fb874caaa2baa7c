"""Import alias: the package directory is named ``text-to-speech-tts-onnx_b200`` (not a valid
Python identifier), so ``import b200tts`` loads it under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "text-to-speech-tts-onnx_b200")
_spec = importlib.util.spec_from_file_location(
    "b200tts", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["b200tts"] = _mod
_spec.loader.exec_module(_mod)
