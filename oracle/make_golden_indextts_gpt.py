"""TEST INFRASTRUCTURE ONLY. Build-container only (needs /root/reference).

Golden vectors for the IndexTTS GPT-2 decode path: the reference's own IndexTTS_B/C/D/E wrapper classes (compiled from
IndexTTS/Export_IndexTTS.py where it lies) around a Hugging Face GPT2Model with synthetic weights, driven by the loop of
IndexTTS/Inference_IndexTTS_ONNX.py:726-781 (restated below with torch tensors in place of OrtValues), on the reduced
configuration config.INDEXTTS_GPT_SMALL. Writes tests/golden/indextts_gpt_ref.npz.

    python -m oracle.make_golden_indextts_gpt
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200tts  # noqa: F401,E402
from b200tts import config, synth  # noqa: E402
from oracle import ref_harness  # noqa: E402

SEED_W, SEED_IN, N_TEXT, MAX_NEW = 555, 31, 12, 40


@torch.inference_mode()
def reference_loop(cfg, sd, conds, text_ids, max_new):
    B, C, D, E = ref_harness.build_indextts_gpt(sd, cfg)
    L, H, hd = cfg.layers, cfg.heads, cfg.head_dim
    text_h = B(torch.from_numpy(text_ids))
    gpt_h, gen_len = C(torch.tensor([[cfg.start_mel]], dtype=torch.int32), torch.tensor([0], dtype=torch.int64))
    gpt_h, concat_len = D(torch.from_numpy(conds), text_h, gpt_h)
    concat_h = gpt_h.clone()
    limit = min(cfg.max_generate - int(concat_len), max_new)
    keys = [torch.zeros((H, hd, 0)) for _ in range(L)]
    vals = [torch.zeros((H, 0, hd)) for _ in range(L)]
    hist = torch.tensor([0], dtype=torch.int64)
    pen = torch.ones((1, cfg.mel_codes))
    ids_len = concat_len.view(1).long()
    mask = torch.tensor([1], dtype=torch.int8)
    ids, hid = [], []
    reset, n = 0, 0
    while n < limit:
        out = E(*keys, *vals, hist, pen, ids_len, gpt_h, mask)
        keys, vals = list(out[:L]), list(out[L:2 * L])
        hist, last, mid = out[2 * L], out[2 * L + 1], out[2 * L + 2]
        tok = int(mid.view(-1)[0])
        ids.append(tok)
        hid.append(last.view(-1).clone())
        n += 1
        if tok == cfg.stop_mel:
            break
        if n < 2:
            mask = torch.tensor([0], dtype=torch.int8)
            ids_len = torch.tensor([1], dtype=torch.int64)
        pen = pen.clone()
        pen[:, tok] = cfg.repeat_penalty
        if n > cfg.penalty_range and ids[reset] != tok:
            pen[:, ids[reset]] = 1.0
            reset += 1
        gpt_h, gen_len = C(mid, gen_len)
    return dict(text_hidden=text_h.numpy(), concat_hidden=concat_h.numpy(), ids=np.asarray(ids, dtype=np.int32),
                hidden=torch.stack(hid).numpy(), penalty=pen.numpy(), key0=keys[0].numpy(), value0=vals[0].numpy())


def main():
    cfg = config.INDEXTTS_GPT_SMALL
    sd = synth.igpt_state(SEED_W, cfg)
    conds, text_ids = synth.igpt_inputs(SEED_IN, N_TEXT, cfg)
    g = reference_loop(cfg, sd, conds, text_ids, MAX_NEW)
    out = os.path.join(ROOT, "tests", "golden", "indextts_gpt_ref.npz")
    np.savez_compressed(out, seed_w=SEED_W, seed_in=SEED_IN, n_text=N_TEXT, max_new=MAX_NEW,
                        **{k: (v.astype(np.float16) if k in ("key0", "value0") else v) for k, v in g.items()})
    print("wrote", out, {k: v.shape for k, v in g.items()}, "ids", g["ids"][:16])


if __name__ == "__main__":
    main()
