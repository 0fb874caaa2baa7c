"""Hyper-parameters of the two graphs on the hot path.

The values the reference pins itself are cited; the rest come from the un-vendored upstream
configs and are confirmed by parameter count.
"""
from dataclasses import dataclass, field
from typing import List


@dataclass(frozen=True)
class BigVGANConfig:
    """bigvgan_v2_24khz_100band_256x (reference: BigVGAN/Export_BigVGAN.py:9, bigvgan.py:282-357)."""
    num_mels: int = 100
    upsample_initial_channel: int = 1536
    upsample_rates: tuple = (4, 4, 2, 2, 2, 2)
    upsample_kernel_sizes: tuple = (8, 8, 4, 4, 4, 4)
    resblock_kernel_sizes: tuple = (3, 7, 11)
    resblock_dilation_sizes: tuple = ((1, 3, 5), (1, 3, 5), (1, 3, 5))
    aa_taps: int = 12            # act.py:12-15
    post_pad: int = 15           # bigvgan.py:370,381-382 (the index -1 tables)
    sample_rate: int = 24000

    @property
    def hop(self) -> int:
        h = 1
        for r in self.upsample_rates:
            h *= r
        return h

    def stage_channels(self) -> List[int]:
        return [self.upsample_initial_channel // (2 ** (i + 1)) for i in range(len(self.upsample_rates))]

    def out_samples(self, frames: int) -> int:
        return frames * self.hop + 2 * self.post_pad


@dataclass(frozen=True)
class F5Config:
    """F5TTS_v1_Base + vocos-mel-24khz (reference: F5_TTS/Export_F5.py:44-65)."""
    dim: int = 1024
    depth: int = 22
    heads: int = 16
    head_dim: int = 64
    ff_mult: int = 2
    text_dim: int = 512
    text_conv_layers: int = 4
    vocab: int = 2545            # text_num_embeds; embedding table has vocab + 1 rows
    n_mels: int = 100
    nfft: int = 1024
    hop: int = 256
    sample_rate: int = 24000
    max_frames: int = 4096       # MAX_SIGNAL_LENGTH, Export_F5.py:59
    nfe: int = 32                # NFE_STEP -> 31 Euler steps
    cfg_strength: float = 2.0
    sway: float = -1.0
    convpos_kernel: int = 31
    convpos_groups: int = 16
    # vocos-mel-24khz
    vocos_dim: int = 512
    vocos_inter: int = 1536
    vocos_layers: int = 8


@dataclass(frozen=True)
class IndexTTSVocoderConfig(BigVGANConfig):
    """The BigVGAN inside IndexTTS_F (reference: IndexTTS/Export_IndexTTS.py:292-314, IndexTTS/modeling_modified/models.py:
    130-250). Stage channels 768..24 are pinned in-repo (IndexTTS/modeling_modified/filter.py:85); gpt_dim, rates and kernel
    sizes come from the un-vendored index-tts config.yaml (rates [4,4,4,4,2,2] -> x1024; kernel sizes [8,8,4,4,4,4], i.e.
    kernel = 2*stride for the first two upsamplers and kernel = stride afterwards -- the engine accepts either per stage)."""
    num_mels: int = 1280         # conv_pre input width = gpt_dim (the GPT latent), not a mel count
    upsample_rates: tuple = (4, 4, 4, 4, 2, 2)
    upsample_kernel_sizes: tuple = (8, 8, 4, 4, 4, 4)
    ln_eps: float = 1e-5         # gpt.final_norm = nn.LayerNorm(gpt_dim)

    @property
    def gpt_dim(self) -> int:
        return self.num_mels


BIGVGAN = BigVGANConfig()
F5 = F5Config()
INDEXTTS_VOCODER = IndexTTSVocoderConfig()


@dataclass(frozen=True)
class IndexTTSGPTConfig:
    """The GPT-2 acoustic model of IndexTTS as graphs B-E run it (reference: IndexTTS/Export_IndexTTS.py:203-289, host loop
    IndexTTS/Inference_IndexTTS_ONNX.py:719-781). In-repo pins: start/stop text ids 0/1 (:208-209), start mel id 8192
    (Inference:673), stop mel id 8193 (:36), MAX_GENERATE_LENGTH 800 (:37), REPEAT_PENALITY 0.7 (:38), PENALITY_RANGE 10
    (:39), q/k pre-scale head_dim^-0.25 (:250-255). model_dim 1280 / 24 layers / 20 heads / 12000 text tokens / 8194 mel
    codes / 32 conditioning latents come from the un-vendored index-tts config.yaml; the position-table sizes are the
    smallest that cover MAX_GENERATE_LENGTH."""
    dim: int = 1280
    layers: int = 24
    heads: int = 20
    head_dim: int = 64
    text_vocab: int = 12000
    mel_codes: int = 8194
    start_text: int = 0
    stop_text: int = 1
    start_mel: int = 8192
    stop_mel: int = 8193
    text_pos: int = 604
    mel_pos: int = 803
    cond_rows: int = 32
    ln_eps: float = 1e-5
    max_generate: int = 800
    repeat_penalty: float = 0.7
    penalty_range: int = 10

    @property
    def ff(self) -> int:
        return 4 * self.dim


INDEXTTS_GPT = IndexTTSGPTConfig()
# reduced copy for the CPU-side golden vectors and quick parity tests (same arithmetic, 3 layers of 8 heads)
INDEXTTS_GPT_SMALL = IndexTTSGPTConfig(dim=512, layers=3, heads=8, text_vocab=300, mel_codes=130, start_mel=128, stop_mel=129,
                                       text_pos=64, mel_pos=100, cond_rows=8, max_generate=96)
