"""Seeded synthetic checkpoints and inputs (no real weights exist offline).

State dicts use the *reference's* ``state_dict`` names (weight-norm already removed, i.e. what
``BigVGAN.from_pretrained(...).remove_weight_norm()`` / ``load_checkpoint(use_ema=True)`` hand
to the export wrappers), so the same dict loads into the reference modules (oracle/ref_harness.py),
the CPU restatement (oracle/*.py) and the engine's converter (weights.py).

numpy's PCG64 ``default_rng`` is used everywhere: its stream is stable across numpy versions,
unlike ``torch.randn``.
"""
import numpy as np

from .config import BIGVGAN, F5, INDEXTTS_VOCODER, BigVGANConfig, F5Config, IndexTTSVocoderConfig


def _n(rng, shape, std):
    return (rng.standard_normal(shape, dtype=np.float32) * np.float32(std)).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# BigVGAN
# ----------------------------------------------------------------------------------------------
def bigvgan_state(seed: int = 1234, cfg: BigVGANConfig = BIGVGAN) -> dict:
    """Synthetic BigVGAN generator weights, scaled so the PCM RMS is ~0.1 full scale."""
    rng = np.random.default_rng(seed)
    sd = {}
    c0 = cfg.upsample_initial_channel
    sd["conv_pre.weight"] = _n(rng, (c0, cfg.num_mels, 7), 0.5 / np.sqrt(cfg.num_mels * 7))
    sd["conv_pre.bias"] = _n(rng, (c0,), 0.05)
    nk = len(cfg.resblock_kernel_sizes)
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        cin, cout = c0 // (2 ** i), c0 // (2 ** (i + 1))
        # ConvTranspose1d weight is (in, out, k); every output sample sees k/u taps of cin channels
        sd[f"ups.{i}.0.weight"] = _n(rng, (cin, cout, k), 1.0 / np.sqrt(cin * k / u))
        sd[f"ups.{i}.0.bias"] = _n(rng, (cout,), 0.05)
        for j, (rk, dil) in enumerate(zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)):
            r = i * nk + j
            for m in range(len(dil)):
                sd[f"resblocks.{r}.convs1.{m}.weight"] = _n(rng, (cout, cout, rk), 0.8 / np.sqrt(cout * rk))
                sd[f"resblocks.{r}.convs1.{m}.bias"] = _n(rng, (cout,), 0.05)
                sd[f"resblocks.{r}.convs2.{m}.weight"] = _n(rng, (cout, cout, rk), 0.6 / np.sqrt(cout * rk))
                sd[f"resblocks.{r}.convs2.{m}.bias"] = _n(rng, (cout,), 0.05)
            for a in range(2 * len(dil)):
                sd[f"resblocks.{r}.activations.{a}.act.alpha"] = _n(rng, (cout,), 0.4)
                sd[f"resblocks.{r}.activations.{a}.act.beta"] = _n(rng, (cout,), 0.4)
    cl = cfg.stage_channels()[-1]
    sd["activation_post.act.alpha"] = _n(rng, (cl,), 0.4)
    sd["activation_post.act.beta"] = _n(rng, (cl,), 0.4)
    sd["conv_post.weight"] = _n(rng, (1, cl, 7), 0.03 / np.sqrt(cl * 7))
    return sd


def ivgan_state(seed: int = 777, cfg: IndexTTSVocoderConfig = INDEXTTS_VOCODER) -> dict:
    """Synthetic weights of the IndexTTS_F vocoder: the BigVGAN generator (reference names of
    IndexTTS/modeling_modified/models.py:130-250, weight norm removed) with a conv_post bias, plus gpt.final_norm."""
    sd = bigvgan_state(seed, cfg)
    rng = np.random.default_rng(seed + 1)
    sd["conv_post.bias"] = _n(rng, (1,), 0.01)
    sd["final_norm.weight"] = (1.0 + _n(rng, (cfg.gpt_dim,), 0.1)).astype(np.float32)
    sd["final_norm.bias"] = _n(rng, (cfg.gpt_dim,), 0.05)
    return sd


def ivgan_inputs(seed: int, rows: int, cfg: IndexTTSVocoderConfig = INDEXTTS_VOCODER):
    """IndexTTS_F inputs (Export_IndexTTS.py:497-520): save_bigvgan_conds_0..5 (1,C_i,1), the cond_layer speaker vector
    (1,1536,1) and save_hidden_state (S, gpt_dim) -- the graph drops the last two rows."""
    rng = np.random.default_rng(seed)
    conds = [_n(rng, (1, c, 1), 0.1) for c in cfg.stage_channels()]
    cond_layer = _n(rng, (1, cfg.upsample_initial_channel, 1), 0.1)
    hidden = (3.0 * rng.standard_normal((rows, cfg.gpt_dim), dtype=np.float32) + 0.5).astype(np.float32)
    return conds, cond_layer, hidden


def igpt_state(seed: int = 555, cfg=None) -> dict:
    """Synthetic weights of the IndexTTS GPT-2 acoustic model under the names graphs B-E read them by (reference:
    IndexTTS/Export_IndexTTS.py:203-289; transformer blocks are Hugging Face GPT2Block: Conv1D weights are (in, out))."""
    from .config import INDEXTTS_GPT
    cfg = cfg or INDEXTTS_GPT
    rng = np.random.default_rng(seed)
    D, FF = cfg.dim, cfg.ff
    sd = {
        "text_embedding.weight": _n(rng, (cfg.text_vocab, D), 0.7),
        "text_pos_embedding.emb.weight": _n(rng, (cfg.text_pos, D), 0.3),
        "mel_embedding.weight": _n(rng, (cfg.mel_codes, D), 0.7),
        "mel_pos_embedding.emb.weight": _n(rng, (cfg.mel_pos, D), 0.3),
    }
    for i in range(cfg.layers):
        p = f"h.{i}."
        sd[p + "ln_1.weight"] = (1.0 + _n(rng, (D,), 0.1)).astype(np.float32)
        sd[p + "ln_1.bias"] = _n(rng, (D,), 0.05)
        sd[p + "attn.c_attn.weight"] = _n(rng, (D, 3 * D), 1.6 / np.sqrt(D))
        sd[p + "attn.c_attn.bias"] = _n(rng, (3 * D,), 0.05)
        sd[p + "attn.c_proj.weight"] = _n(rng, (D, D), 0.5 / np.sqrt(D))
        sd[p + "attn.c_proj.bias"] = _n(rng, (D,), 0.02)
        sd[p + "ln_2.weight"] = (1.0 + _n(rng, (D,), 0.1)).astype(np.float32)
        sd[p + "ln_2.bias"] = _n(rng, (D,), 0.05)
        sd[p + "mlp.c_fc.weight"] = _n(rng, (D, FF), 1.0 / np.sqrt(D))
        sd[p + "mlp.c_fc.bias"] = _n(rng, (FF,), 0.05)
        sd[p + "mlp.c_proj.weight"] = _n(rng, (FF, D), 0.5 / np.sqrt(FF))
        sd[p + "mlp.c_proj.bias"] = _n(rng, (D,), 0.02)
    for nm in ("ln_f", "final_norm"):
        sd[nm + ".weight"] = (1.0 + _n(rng, (D,), 0.1)).astype(np.float32)
        sd[nm + ".bias"] = _n(rng, (D,), 0.05)
    sd["mel_head.weight"] = _n(rng, (cfg.mel_codes, D), 2.0 / np.sqrt(D))
    sd["mel_head.bias"] = _n(rng, (cfg.mel_codes,), 0.1)
    return sd


def igpt_inputs(seed: int, n_text: int, cfg=None):
    """-> conds_latent (1, cond_rows, D) f32 (graph A's output, Export_IndexTTS.py:200), text_ids (1, n_text) i32."""
    from .config import INDEXTTS_GPT
    cfg = cfg or INDEXTTS_GPT
    rng = np.random.default_rng(seed)
    conds = _n(rng, (1, cfg.cond_rows, cfg.dim), 0.8)
    ids = rng.integers(2, cfg.text_vocab, size=(1, n_text)).astype(np.int32)
    return conds, ids


def bigvgan_mel(seed: int, batch: int, frames: int, cfg: BigVGANConfig = BIGVGAN) -> np.ndarray:
    """Log-mel shaped input (B, n_mels, T): 2*randn-4 clipped to [ln 1e-5, 3]."""
    rng = np.random.default_rng(seed)
    m = 2.0 * rng.standard_normal((batch, cfg.num_mels, frames), dtype=np.float32) - 4.0
    return np.clip(m, np.log(1e-5), 3.0).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# F5-TTS: DiT + text embedding, Vocos
# ----------------------------------------------------------------------------------------------
def f5_dit_state(seed: int = 4321, cfg: F5Config = F5) -> dict:
    """Synthetic F5 DiT weights (reference names: F5_TTS/modeling_modified/F5/dit.py).

    The tensors upstream zero-initialises (AdaLN linears, proj_out; dit.py:156-166) are
    re-randomised, otherwise the DiT output is identically zero."""
    rng = np.random.default_rng(seed)
    D, H = cfg.dim, cfg.dim * cfg.ff_mult
    sd = {}
    sd["time_embed.time_mlp.0.weight"] = _n(rng, (D, 256), 1.0 / 16.0)
    sd["time_embed.time_mlp.0.bias"] = _n(rng, (D,), 0.02)
    sd["time_embed.time_mlp.2.weight"] = _n(rng, (D, D), 1.0 / 32.0)
    sd["time_embed.time_mlp.2.bias"] = _n(rng, (D,), 0.02)
    td = cfg.text_dim
    sd["text_embed.text_embed.weight"] = _n(rng, (cfg.vocab + 1, td), 1.0)
    for i in range(cfg.text_conv_layers):
        p = f"text_embed.text_blocks.{i}."
        sd[p + "dwconv.weight"] = _n(rng, (td, 1, 7), 1.0 / np.sqrt(7.0))
        sd[p + "dwconv.bias"] = _n(rng, (td,), 0.02)
        sd[p + "norm.weight"] = (1.0 + _n(rng, (td,), 0.1)).astype(np.float32)
        sd[p + "norm.bias"] = _n(rng, (td,), 0.05)
        sd[p + "pwconv1.weight"] = _n(rng, (2 * td, td), 1.0 / np.sqrt(td))
        sd[p + "pwconv1.bias"] = _n(rng, (2 * td,), 0.02)
        sd[p + "grn.gamma"] = _n(rng, (1, 1, 2 * td), 0.2)
        sd[p + "grn.beta"] = _n(rng, (1, 1, 2 * td), 0.05)
        sd[p + "pwconv2.weight"] = _n(rng, (td, 2 * td), 0.5 / np.sqrt(2 * td))
        sd[p + "pwconv2.bias"] = _n(rng, (td,), 0.02)
    kin = 2 * cfg.n_mels + td
    sd["input_embed.proj.weight"] = _n(rng, (D, kin), 1.0 / np.sqrt(kin))
    sd["input_embed.proj.bias"] = _n(rng, (D,), 0.02)
    gc = D // cfg.convpos_groups
    for n in (0, 2):
        sd[f"input_embed.conv_pos_embed.conv1d.{n}.weight"] = _n(
            rng, (D, gc, cfg.convpos_kernel), 1.0 / np.sqrt(gc * cfg.convpos_kernel))
        sd[f"input_embed.conv_pos_embed.conv1d.{n}.bias"] = _n(rng, (D,), 0.02)
    for i in range(cfg.depth):
        p = f"transformer_blocks.{i}."
        sd[p + "attn_norm.linear.weight"] = _n(rng, (6 * D, D), 0.015)
        sd[p + "attn_norm.linear.bias"] = _n(rng, (6 * D,), 0.05)
        for nm in ("to_q", "to_k", "to_v", "to_out.0"):
            sd[p + f"attn.{nm}.weight"] = _n(rng, (D, D), 1.0 / np.sqrt(D))
            sd[p + f"attn.{nm}.bias"] = _n(rng, (D,), 0.02)
        sd[p + "ff.ff.0.0.weight"] = _n(rng, (H, D), 1.0 / np.sqrt(D))
        sd[p + "ff.ff.0.0.bias"] = _n(rng, (H,), 0.02)
        sd[p + "ff.ff.2.weight"] = _n(rng, (D, H), 1.0 / np.sqrt(H))
        sd[p + "ff.ff.2.bias"] = _n(rng, (D,), 0.02)
    sd["norm_out.linear.weight"] = _n(rng, (2 * D, D), 0.015)
    sd["norm_out.linear.bias"] = _n(rng, (2 * D,), 0.05)
    sd["proj_out.weight"] = _n(rng, (cfg.n_mels, D), 1.0 / np.sqrt(D))
    sd["proj_out.bias"] = _n(rng, (cfg.n_mels,), 0.02)
    return sd


def vocos_state(seed: int = 2468, cfg: F5Config = F5) -> dict:
    """Synthetic vocos-mel-24khz weights with the upstream (un-folded) names and shapes;
    the Export_F5.py:390-402 folding is applied by the consumers."""
    rng = np.random.default_rng(seed)
    C, I = cfg.vocos_dim, cfg.vocos_inter
    sd = {}
    sd["backbone.embed.weight"] = _n(rng, (C, cfg.n_mels, 7), 1.0 / np.sqrt(cfg.n_mels * 7))
    sd["backbone.embed.bias"] = _n(rng, (C,), 0.02)
    sd["backbone.norm.weight"] = (1.0 + _n(rng, (C,), 0.1)).astype(np.float32)
    sd["backbone.norm.bias"] = _n(rng, (C,), 0.05)
    for i in range(cfg.vocos_layers):
        p = f"backbone.convnext.{i}."
        sd[p + "dwconv.weight"] = _n(rng, (C, 1, 7), 1.0 / np.sqrt(7.0))
        sd[p + "dwconv.bias"] = _n(rng, (C,), 0.02)
        sd[p + "norm.weight"] = (1.0 + _n(rng, (C,), 0.1)).astype(np.float32)
        sd[p + "norm.bias"] = _n(rng, (C,), 0.05)
        sd[p + "pwconv1.weight"] = _n(rng, (I, C), 1.0 / np.sqrt(C))
        sd[p + "pwconv1.bias"] = _n(rng, (I,), 0.02)
        sd[p + "pwconv2.weight"] = _n(rng, (C, I), 1.0 / np.sqrt(I))
        sd[p + "pwconv2.bias"] = _n(rng, (C,), 0.02)
        sd[p + "gamma"] = (0.3 * (1.0 + _n(rng, (C,), 0.1))).astype(np.float32)
    sd["backbone.final_layer_norm.weight"] = (1.0 + _n(rng, (C,), 0.1)).astype(np.float32)
    sd["backbone.final_layer_norm.bias"] = _n(rng, (C,), 0.05)
    sd["head.out.weight"] = _n(rng, (cfg.nfft + 2, C), 0.7 / np.sqrt(C))
    b = _n(rng, (cfg.nfft + 2,), 0.05)
    b[: cfg.nfft // 2 + 1] += 1.0      # log-magnitude rows: PCM RMS ~0.1, exp(.) mostly under the clip at 100
    sd["head.out.bias"] = b
    return sd


def f5_inputs(seed: int, audio_len: int = 144000, n_text: int = 150, cfg: F5Config = F5):
    """Config-3 style inputs: int16 audio (1,1,L), int32 text ids (1,n), int64 [N],
    and the Euler start noise (1,N,n_mels) that the reference draws inside graph A."""
    rng = np.random.default_rng(seed)
    audio = np.clip(3000.0 * rng.standard_normal(audio_len), -32768, 32767).astype(np.int16).reshape(1, 1, -1)
    text_ids = rng.integers(0, cfg.vocab, size=(1, n_text), dtype=np.int32)
    ref_len = audio_len // cfg.hop + 1
    max_duration = np.array([2 * ref_len], dtype=np.int64)     # ref_text_len == gen_text_len
    noise = np.random.default_rng(seed + 9527).standard_normal(
        (1, int(max_duration[0]), cfg.n_mels), dtype=np.float32)
    return audio, text_ids, max_duration, noise
