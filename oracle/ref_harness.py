"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). Build-container only: needs /root/reference.

Loads the reference's own ``modeling_modified`` nn.Module sources *where they lie* under
/root/reference (nothing is copied) through small stubs for the un-installed third-party imports,
so that oracle/*_ref.py can be pinned against the real thing and tests/golden/ can be generated
(oracle/make_golden.py). Never used at run time on the GPU box (the reference is not there).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("B200TTS_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "BigVGAN", "modeling_modified"))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


# ----------------------------------------------------------------------------------------------
# BigVGAN
# ----------------------------------------------------------------------------------------------
class _SnakeBeta(torch.nn.Module):
    """Upstream NVIDIA/BigVGAN activations.SnakeBeta (not vendored by the reference); formula as in
    Qwen_TTS/modeling_modified/modeling_qwen3_tts_tokenizer_v2.py:665-684."""

    def __init__(self, in_features, alpha=1.0, alpha_trainable=True, alpha_logscale=False):
        super().__init__()
        self.alpha_logscale = alpha_logscale
        init = torch.zeros(in_features) if alpha_logscale else torch.ones(in_features)
        self.alpha = torch.nn.Parameter(init * alpha if not alpha_logscale else init.clone())
        self.beta = torch.nn.Parameter(init * alpha if not alpha_logscale else init.clone())
        self.no_div_by_zero = 1e-9

    def forward(self, x):
        alpha = self.alpha.unsqueeze(0).unsqueeze(-1)
        beta = self.beta.unsqueeze(0).unsqueeze(-1)
        if self.alpha_logscale:
            alpha = torch.exp(alpha)
            beta = torch.exp(beta)
        return x + (1.0 / (beta + self.no_div_by_zero)) * torch.pow(torch.sin(x * alpha), 2)


class _AttrDict(dict):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.__dict__ = self


def load_bigvgan_module():
    d = os.path.join(REF, "BigVGAN", "modeling_modified")
    _stub("activations", SnakeBeta=_SnakeBeta, Snake=_SnakeBeta)
    _stub("utils", init_weights=lambda m, mean=0.0, std=0.01: None,
          get_padding=lambda k, dil=1: int((k * dil - dil) / 2))
    _stub("env", AttrDict=_AttrDict)
    _stub("alias_free_activation")
    _stub("alias_free_activation.torch")
    _load("alias_free_activation.torch.filter", os.path.join(d, "filter.py"))
    _load("alias_free_activation.torch.resample", os.path.join(d, "resample.py"))
    _load("alias_free_activation.torch.act", os.path.join(d, "act.py"))
    return _load("ref_bigvgan", os.path.join(d, "bigvgan.py"))


def build_bigvgan(sd: dict, cfg):
    """Reference BigVGAN generator filled from a synthetic state dict, wrapped exactly as
    BigVGAN/Export_BigVGAN.py:37-57 does (weight norm removed, eval, float, tanh forced)."""
    mod = load_bigvgan_module()
    h = _AttrDict(
        num_mels=cfg.num_mels, upsample_rates=list(cfg.upsample_rates),
        upsample_kernel_sizes=list(cfg.upsample_kernel_sizes),
        upsample_initial_channel=cfg.upsample_initial_channel, resblock="1",
        resblock_kernel_sizes=list(cfg.resblock_kernel_sizes),
        resblock_dilation_sizes=[list(d) for d in cfg.resblock_dilation_sizes],
        activation="snakebeta", snake_logscale=True, use_bias_at_final=False, use_tanh_at_final=False)
    model = mod.BigVGAN(h, use_cuda_kernel=False)
    model.remove_weight_norm()
    missing, unexpected = model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    bad = [m for m in missing if not m.endswith("filter")]
    assert not bad and not unexpected, (bad, unexpected)
    model = model.eval().float()

    class BIGVGAN(torch.nn.Module):          # restated from Export_BigVGAN.py:37-49 (script is not importable)
        def __init__(self, bigvgan, use_tanh):
            super().__init__()
            self.bigvgan = bigvgan
            self.bigvgan.use_tanh_at_final = use_tanh
            self.use_tanh = use_tanh

        def forward(self, mel_features):
            w = self.bigvgan(mel_features) * 32767.0
            if self.use_tanh:
                w = w.clamp(min=-32768.0, max=32767.0)
            return w.to(torch.int16)

    return BIGVGAN(model, True)


# ----------------------------------------------------------------------------------------------
# IndexTTS_F: the vocoder half of IndexTTS
# ----------------------------------------------------------------------------------------------
def load_indextts_bigvgan_module():
    d = os.path.join(REF, "IndexTTS", "modeling_modified")

    class _ECAPA(torch.nn.Module):            # conditioning branch: not part of graph F (its outputs are graph inputs)
        def __init__(self, *a, **k):
            super().__init__()

    _stub("indextts")
    _stub("indextts.BigVGAN")
    _stub("indextts.BigVGAN.activations", SnakeBeta=_SnakeBeta, Snake=_SnakeBeta)
    _stub("indextts.BigVGAN.ECAPA_TDNN", ECAPA_TDNN=_ECAPA)
    _stub("indextts.BigVGAN.utils", init_weights=lambda m, mean=0.0, std=0.01: None,
          get_padding=lambda k, dil=1: int((k * dil - dil) / 2))
    pkg = _stub("indextts.BigVGAN.alias_free_torch")
    pkg.__path__ = [d]                        # so that act.py / resample.py resolve their relative imports
    _load("indextts.BigVGAN.alias_free_torch.filter", os.path.join(d, "filter.py"))
    _load("indextts.BigVGAN.alias_free_torch.resample", os.path.join(d, "resample.py"))
    act = _load("indextts.BigVGAN.alias_free_torch.act", os.path.join(d, "act.py"))
    pkg.Activation1d = act.Activation1d
    return _load("ref_indextts_models", os.path.join(d, "models.py"))


def build_indextts_f(sd: dict, cfg):
    """The reference IndexTTS BigVGAN (IndexTTS/modeling_modified/models.py) filled from a synthetic state dict, plus
    gpt.final_norm, wrapped as IndexTTS/Export_IndexTTS.py:292-314 does (the script itself is not importable: it loads
    checkpoints at import time)."""
    mod = load_indextts_bigvgan_module()
    h = _AttrDict(
        num_mels=100, gpt_dim=cfg.gpt_dim, upsample_rates=list(cfg.upsample_rates),
        upsample_kernel_sizes=list(cfg.upsample_kernel_sizes), upsample_initial_channel=cfg.upsample_initial_channel,
        resblock="1", resblock_kernel_sizes=list(cfg.resblock_kernel_sizes),
        resblock_dilation_sizes=[list(d) for d in cfg.resblock_dilation_sizes], activation="snakebeta", snake_logscale=True,
        feat_upsample=False, cond_d_vector_in_each_upsampling_layer=True, speaker_embedding_dim=512)
    model = mod.BigVGAN(h, use_cuda_kernel=False)
    model.remove_weight_norm()
    gen = {k: torch.from_numpy(v) for k, v in sd.items() if not k.startswith("final_norm.")}
    missing, unexpected = model.load_state_dict(gen, strict=False)
    bad = [m for m in missing if not m.endswith("filter") and not m.startswith(("cond_layer.", "conds.", "speaker_encoder."))]
    assert not bad and not unexpected, (bad, unexpected)
    model = model.eval().float()
    final_norm = torch.nn.LayerNorm(cfg.gpt_dim, eps=cfg.ln_eps)
    final_norm.load_state_dict({"weight": torch.from_numpy(sd["final_norm.weight"]), "bias": torch.from_numpy(sd["final_norm.bias"])})

    class IndexTTS_F(torch.nn.Module):        # restated from Export_IndexTTS.py:292-314
        def __init__(self):
            super().__init__()
            self.bigvgan, self.final_norm = model, final_norm.eval()
            self.inv_num_kernels = float(1.0 / model.num_kernels)

        def forward(self, *all_inputs):
            latent = self.final_norm(all_inputs[-1][:-2].unsqueeze(0))
            latent = self.bigvgan.conv_pre(latent.transpose(1, 2)) + all_inputs[-2]
            for i in range(self.bigvgan.num_upsamples):
                for i_up in range(len(self.bigvgan.ups[i])):
                    latent = self.bigvgan.ups[i][i_up](latent)
                if self.bigvgan.cond_in_each_up_layer:
                    latent = latent + all_inputs[i]
                x = self.bigvgan.resblocks[i * self.bigvgan.num_kernels](latent, i)
                for j in range(1, self.bigvgan.num_kernels):
                    x = x + self.bigvgan.resblocks[i * self.bigvgan.num_kernels + j](latent, i)
                latent = x * self.inv_num_kernels
            latent = self.bigvgan.conv_post(self.bigvgan.activation_post(latent, -1))
            generated_wav = torch.tanh(latent)
            return (generated_wav.clamp(min=-1.0, max=1.0) * 32767.0).to(torch.int16)

    return IndexTTS_F()


def build_indextts_gpt(sd: dict, cfg):
    """The reference's IndexTTS_B / _C / _D / _E wrapper classes, compiled from Export_IndexTTS.py where it lies (the script
    loads checkpoints at import time, so the class definitions are extracted and exec'd; nothing is copied), driving a Hugging
    Face GPT2Model -- the class index-tts (un-vendored) builds its inference_model from -- filled with a synthetic state dict.
    -> (B, C, D, E) modules. B and D take a Python int for a tensor shape (they were written for the tracer), so they are
    returned as callables that run the module under torch.jit.trace."""
    import ast
    from transformers import GPT2Config, GPT2Model
    path = os.path.join(REF, "IndexTTS", "Export_IndexTTS.py")
    tree = ast.parse(open(path, encoding="utf-8").read())
    wanted = {"IndexTTS_B", "IndexTTS_C", "IndexTTS_D", "IndexTTS_E"}
    body = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in wanted]
    assert {n.name for n in body} == wanted
    ns = {"torch": torch}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)

    t = lambda k: torch.from_numpy(np.ascontiguousarray(sd[k])).float()
    D = cfg.dim
    gcfg = GPT2Config(vocab_size=8, n_positions=8, n_embd=D, n_layer=cfg.layers, n_head=cfg.heads,
                      activation_function="gelu_new", layer_norm_epsilon=cfg.ln_eps, resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
    tr = GPT2Model(gcfg)
    own = {k: t(k) for k in sd if k.startswith(("h.", "ln_f."))}
    missing, unexpected = tr.load_state_dict(own, strict=False)
    assert not unexpected and all(m.startswith(("wte.", "wpe.")) or m.endswith((".attn.bias", ".attn.masked_bias")) for m in missing), \
        (missing, unexpected)

    def emb(k):
        e = torch.nn.Embedding(sd[k].shape[0], D)
        e.weight.data = t(k)
        return e

    class _Pos(torch.nn.Module):
        def __init__(self, k):
            super().__init__()
            self.emb = emb(k)

    class _Inf(torch.nn.Module):              # GPT2InferenceModel's attribute names as Export_IndexTTS.py uses them
        def __init__(self):
            super().__init__()
            self.transformer = tr
            self.embeddings = emb("mel_embedding.weight")
            self.text_pos_embedding = _Pos("mel_pos_embedding.emb.weight")       # the MEL position table (index-tts naming)
            fn = torch.nn.LayerNorm(D, eps=cfg.ln_eps)
            fn.weight.data, fn.bias.data = t("final_norm.weight"), t("final_norm.bias")
            head = torch.nn.Linear(D, cfg.mel_codes)
            head.weight.data, head.bias.data = t("mel_head.weight"), t("mel_head.bias")
            self.lm_head = torch.nn.Sequential(fn, head)

    class _GPT(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.text_embedding = emb("text_embedding.weight")
            self.text_pos_embedding = _Pos("text_pos_embedding.emb.weight")
            self.inference_model = _Inf()

    class _Index:
        pass

    idx = _Index()
    idx.gpt = _GPT().eval()
    with torch.inference_mode():
        B = ns["IndexTTS_B"](idx).eval()
        C = ns["IndexTTS_C"](idx).eval()
        Dm = ns["IndexTTS_D"]().eval()
        E = ns["IndexTTS_E"](idx, cfg.layers, cfg.max_generate).eval()
    traced = lambda mod: (lambda *a: torch.jit.trace(mod, a, check_trace=False)(*a))
    return traced(B), C, traced(Dm), E


# ----------------------------------------------------------------------------------------------
# F5-TTS: DiT, STFT_Process, Vocos and the three export wrappers
# ----------------------------------------------------------------------------------------------
def load_f5_modules():
    """-> (dit module, STFT_Process class, vocos pretrained module), loaded from the reference tree."""
    d = os.path.join(REF, "F5_TTS")
    mm = os.path.join(d, "modeling_modified")
    _stub("librosa")
    _stub("librosa.filters", mel=lambda *a, **k: None)
    _stub("x_transformers")

    class _Rotary(torch.nn.Module):          # only constructed, never called on the exported path (dit.py:130)
        def __init__(self, dim):
            super().__init__()

    _stub("x_transformers.x_transformers", apply_rotary_pos_emb=lambda *a, **k: None, RotaryEmbedding=_Rotary)
    _stub("f5_tts")
    _stub("f5_tts.model")
    _load("f5_tts.model.modules", os.path.join(mm, "F5", "modules.py"))
    dit = _load("ref_f5_dit", os.path.join(mm, "F5", "dit.py"))
    if "onnxruntime" not in sys.modules:
        _stub("onnxruntime")                 # STFT_Process.py:4 imports it for its own print-only tests
    stft = _load("ref_stft_process", os.path.join(d, "STFT_Process.py"))
    _stub("vocos")
    _stub("vocos.spectral_ops", ISTFT=lambda **k: None, IMDCT=lambda **k: None)
    _stub("vocos.feature_extractors", FeatureExtractor=torch.nn.Module, EncodecFeatures=type("EncodecFeatures", (), {}))
    _load("vocos.modules", os.path.join(mm, "vocos", "modules.py"))
    _load("vocos.models", os.path.join(mm, "vocos", "models.py"))
    _load("vocos.heads", os.path.join(mm, "vocos", "heads.py"))
    voc = _load("vocos.pretrained", os.path.join(mm, "vocos", "pretrained.py"))
    return dit, stft.STFT_Process, voc


def _export_wrappers(namespace):
    """Compile the three wrapper classes straight from the reference's Export_F5.py source (lines 98-203). The
    script itself cannot be imported (top-level file copies and checkpoint loads), and nothing is copied into this
    repo: the class definitions are extracted from the file where it lies and exec'd in ``namespace``."""
    import ast
    path = os.path.join(REF, "F5_TTS", "Export_F5.py")
    tree = ast.parse(open(path, encoding="utf-8").read())
    wanted = {"F5Preprocess", "F5Transformer", "F5Decode"}
    body = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in wanted]
    assert {n.name for n in body} == wanted
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), namespace)
    return namespace


def build_f5(dit_sd: dict, vocos_sd: dict, cfg):
    """Reference F5Preprocess / F5Transformer / F5Decode modules on synthetic checkpoints, transformed exactly as
    Export_F5.py does (Q/K pre-scale :321-333 between exporting A and B, Vocos folding :389-402)."""
    import math
    import torchaudio
    dit_mod, STFT_Process, voc = load_f5_modules()
    ns = _export_wrappers({"torch": torch, "torchaudio": torchaudio, "math": math, "MAX_SIGNAL_LENGTH": cfg.max_frames})

    transformer = dit_mod.DiT(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, ff_mult=cfg.ff_mult, text_dim=cfg.text_dim,
                              conv_layers=cfg.text_conv_layers, text_num_embeds=cfg.vocab, mel_dim=cfg.n_mels)
    missing, unexpected = transformer.load_state_dict({k: torch.from_numpy(v) for k, v in dit_sd.items()}, strict=False)
    assert not unexpected and all("rotary" in m or "freqs_cis" in m for m in missing), (missing, unexpected)
    transformer = transformer.eval().float()

    class _CFM:                       # the wrappers only touch f5_model.transformer (Export_F5.py:101,147-148)
        pass

    f5_model = _CFM()
    f5_model.transformer = transformer
    with torch.inference_mode():
        stft = STFT_Process(model_type="stft_B", n_fft=cfg.nfft, win_length=cfg.nfft, hop_len=cfg.hop, max_frames=0,
                            window_type="hann").eval()
        pre = ns["F5Preprocess"](f5_model, stft, nfft=cfg.nfft, n_mels=cfg.n_mels, sample_rate=cfg.sample_rate,
                                 num_head=cfg.heads, head_dim=cfg.head_dim, target_rms=0.15, use_fp16=False)
        scale = math.pow(cfg.head_dim, -0.25)
        for blk in transformer.transformer_blocks:          # Export_F5.py:329-333
            blk.attn.to_q.weight.data *= scale
            blk.attn.to_q.bias.data *= scale
            blk.attn.to_k.weight.data *= scale
            blk.attn.to_k.bias.data *= scale
        trans = ns["F5Transformer"](f5_model, cfg=cfg.cfg_strength, steps=cfg.nfe, sway_coef=cfg.sway, dtype=torch.float32,
                                    fuse_step=1)

        istft = STFT_Process(model_type="istft_A", n_fft=cfg.nfft, win_length=cfg.nfft, hop_len=cfg.hop,
                             max_frames=cfg.max_frames, window_type="hann").eval()
        import vocos.heads as vh
        import vocos.models as vm
        backbone = vm.VocosBackbone(input_channels=cfg.n_mels, dim=cfg.vocos_dim, intermediate_dim=cfg.vocos_inter,
                                    num_layers=cfg.vocos_layers)
        vh.ISTFT = lambda **k: torch.nn.Identity()
        head = vh.ISTFTHead(dim=cfg.vocos_dim, n_fft=cfg.nfft, hop_length=cfg.hop, padding="center")
        vocos = voc.Vocos(feature_extractor=torch.nn.Identity(), backbone=backbone, head=head)
        missing, unexpected = vocos.load_state_dict({k: torch.from_numpy(v) for k, v in vocos_sd.items()}, strict=False)
        assert not missing and not unexpected, (missing, unexpected)
        vocos.eval()
        # Export_F5.py:390-402, verbatim semantics
        rt = lambda w: torch.sqrt(torch.tensor(w.shape[0], dtype=torch.float32))
        bb = vocos.backbone
        bb.norm.weight.data = (bb.norm.weight.data * rt(bb.norm.weight.data)).view(1, -1, 1)
        bb.norm.bias.data = bb.norm.bias.data.view(1, -1, 1)
        bb.final_layer_norm.weight.data = (bb.final_layer_norm.weight.data * rt(bb.final_layer_norm.weight.data)).view(1, -1, 1)
        bb.final_layer_norm.bias.data = bb.final_layer_norm.bias.data.view(1, -1, 1)
        vocos.head.out.bias.data = vocos.head.out.bias.data.view(1, -1, 1)
        for block in bb.convnext:
            block.norm.weight.data = (block.norm.weight.data * rt(block.norm.weight.data)).view(1, -1, 1)
            block.norm.bias.data = block.norm.bias.data.view(1, -1, 1)
            block.pwconv1.weight.data = block.pwconv1.weight.data.unsqueeze(0)
            block.pwconv1.bias.data = block.pwconv1.bias.data.view(1, -1, 1)
            block.pwconv2.weight.data = (block.gamma.data.unsqueeze(-1) * block.pwconv2.weight.data).unsqueeze(0)
            block.pwconv2.bias.data = (block.gamma.data * block.pwconv2.bias.data).view(1, -1, 1)
        dec = ns["F5Decode"](vocos, istft, target_rms=0.15, use_fp16=False)
    return pre.eval(), trans.eval(), dec.eval()
