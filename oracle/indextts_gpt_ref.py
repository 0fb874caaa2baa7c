"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of the acoustic half of IndexTTS -- graphs B, C, D, E and the greedy decode loop around them -- in PyTorch
fp32 eager. Follows, quirks included:
  IndexTTS/Export_IndexTTS.py:203-214   IndexTTS_B: [start 0, text ids, stop 1] -> text_embedding + text_pos_embedding[:len]
  IndexTTS/Export_IndexTTS.py:217-225   IndexTTS_C: mel_embedding(id) + mel_pos_embedding[gen_len]; gen_len + 1
  IndexTTS/Export_IndexTTS.py:228-235   IndexTTS_D: concat(conds_latent, text rows, first mel row) on the row axis
  IndexTTS/Export_IndexTTS.py:238-262   IndexTTS_E.__init__: HF Conv1D weights transposed, q and k rows (and biases) scaled
                                        by head_dim^-0.25 each, split per head; int8 mask table (1 - tril) * -128
  IndexTTS/Export_IndexTTS.py:264-289   IndexTTS_E.forward: per layer ln_1 -> per-head q/k/v -> K cache kept TRANSPOSED
                                        (H, 64, S), V (H, S, 64) -> softmax(q@k + mask*flag) @ v -> per-head c_proj summed over
                                        heads + bias -> residual; ln_2 -> c_fc -> gelu_new -> c_proj -> residual; ln_f on the last
                                        row; lm_head (= final_norm -> mel_head) * repeat_penality; argmax
  IndexTTS/Inference_IndexTTS_ONNX.py:726-781   host loop: prefill with the mask flag 1, then one row per call with flag 0;
                                        the hidden state of EVERY call is kept (the stop token's too -- graph F drops the
                                        last two rows); penalty[id] = 0.7 after each token, and once more than PENALITY_RANGE
                                        tokens were produced the oldest penalised id is released (set to 1.0) unless it equals
                                        the current one; the penalty vector is never reset between sentences.
The transformer blocks are Hugging Face GPT2Block modules in the reference (index-tts, un-vendored; transformers is
unpinned): LayerNorm eps 1e-5, gelu_new = 0.5x(1 + tanh(sqrt(2/pi)(x + 0.044715 x^3))).
Pinned against the reference's own wrapper classes (extracted from Export_IndexTTS.py where it lies) driving a Hugging Face
GPT2Model by oracle/ref_harness.py::build_indextts_gpt (tests/golden/indextts_gpt_ref.npz).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def _t(sd, k):
    v = sd[k]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))


def fold_layer(sd, i, cfg):
    """Export_IndexTTS.py:249-262 for layer i -> dict of per-head tensors."""
    D, H, hd = cfg.dim, cfg.heads, cfg.head_dim
    w = _t(sd, f"h.{i}.attn.c_attn.weight").float().transpose(0, 1).clone()      # (3D, D)
    b = _t(sd, f"h.{i}.attn.c_attn.bias").float().clone()
    s = float(hd ** -0.25)
    w[:2 * D] *= s
    b[:2 * D] *= s
    out = {}
    for n, lo in (("q", 0), ("k", D), ("v", 2 * D)):
        out[n + "_w"] = w[lo:lo + D].view(H, hd, D).transpose(1, 2).contiguous()  # (H, D, hd)
        out[n + "_b"] = b[lo:lo + D].view(H, 1, hd).contiguous()
    pw = _t(sd, f"h.{i}.attn.c_proj.weight").float().transpose(0, 1)              # (D_out, D_in)
    out["o_w"] = pw.reshape(D, H, hd).permute(1, 2, 0).contiguous()               # (H, hd, D_out)
    out["o_b"] = _t(sd, f"h.{i}.attn.c_proj.bias").float().view(1, 1, -1)
    return out


def gelu_new(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


@torch.inference_mode()
def text_embed(text_ids, sd, cfg):
    """IndexTTS_B: (1, n) int -> (1, n + 2, D)."""
    ids = torch.as_tensor(np.asarray(text_ids)).long().view(1, -1)
    ids = torch.cat([torch.tensor([[cfg.start_text]]), ids, torch.tensor([[cfg.stop_text]])], dim=-1)
    n = ids.shape[-1]
    return F.embedding(ids, _t(sd, "text_embedding.weight").float()) + _t(sd, "text_pos_embedding.emb.weight").float()[:n]


@torch.inference_mode()
def mel_embed(mel_id, gen_len, sd):
    """IndexTTS_C: id (1, 1), gen_len int -> ((1, 1, D), gen_len + 1)."""
    ids = torch.as_tensor(np.asarray(mel_id)).long().view(1, 1)
    h = F.embedding(ids, _t(sd, "mel_embedding.weight").float())
    h = h + _t(sd, "mel_pos_embedding.emb.weight").float()[int(gen_len)].view(1, -1)
    return h, int(gen_len) + 1


@torch.inference_mode()
def step_e(folded, past_k, past_v, penalty, hidden, mask_flag, sd, cfg):
    """IndexTTS_E.forward. past_k[i] (H, hd, S), past_v[i] (H, S, hd), penalty (1, mel_codes), hidden (1, n, D), mask_flag 0|1
    -> (keys, values, last_hidden (1, D), max id (1, 1) int32, logits (1, mel_codes))."""
    n = hidden.shape[1]
    hist = past_k[0].shape[2]
    kv = hist + n
    mask8 = (1 - torch.tril(torch.ones([1, kv, kv], dtype=torch.int8))) * -128
    mask = (mask8[:, :n, :kv] * int(mask_flag)).float()
    D = cfg.dim
    keys, vals = [], []
    h = hidden.float().clone()
    for i in range(cfg.layers):
        f = folded[i]
        x = F.layer_norm(h, (D,), _t(sd, f"h.{i}.ln_1.weight").float(), _t(sd, f"h.{i}.ln_1.bias").float(), cfg.ln_eps)
        q = torch.matmul(x, f["q_w"]) + f["q_b"]
        k = (torch.matmul(x, f["k_w"]) + f["k_b"]).transpose(1, 2)
        v = torch.matmul(x, f["v_w"]) + f["v_b"]
        k = torch.cat((past_k[i], k), dim=2)
        v = torch.cat((past_v[i], v), dim=1)
        keys.append(k)
        vals.append(v)
        a = torch.matmul(torch.softmax(torch.matmul(q, k) + mask, dim=-1), v)
        a = torch.matmul(a, f["o_w"]).sum(dim=0, keepdim=True) + f["o_b"]
        h = h + a
        y = F.layer_norm(h, (D,), _t(sd, f"h.{i}.ln_2.weight").float(), _t(sd, f"h.{i}.ln_2.bias").float(), cfg.ln_eps)
        y = torch.addmm(_t(sd, f"h.{i}.mlp.c_fc.bias").float(), y.view(-1, D), _t(sd, f"h.{i}.mlp.c_fc.weight").float())
        y = gelu_new(y)
        y = torch.addmm(_t(sd, f"h.{i}.mlp.c_proj.bias").float(), y, _t(sd, f"h.{i}.mlp.c_proj.weight").float())
        h = h + y.view(1, n, D)
    last = F.layer_norm(h[:, -1], (D,), _t(sd, "ln_f.weight").float(), _t(sd, "ln_f.bias").float(), cfg.ln_eps)
    z = F.layer_norm(last, (D,), _t(sd, "final_norm.weight").float(), _t(sd, "final_norm.bias").float(), cfg.ln_eps)
    logits = F.linear(z, _t(sd, "mel_head.weight").float(), _t(sd, "mel_head.bias").float()) * penalty
    ids = torch.argmax(logits, dim=-1, keepdim=True).int()
    return keys, vals, last, ids, logits


@torch.inference_mode()
def generate(conds_latent, text_ids, sd, cfg, max_new=None, penalty=None, return_logits=False):
    """Graphs B, C, D and the decode loop of Inference_IndexTTS_ONNX.py:726-781 for one sentence.
    conds_latent (1, R, D). -> (ids (n,) int32, hidden (n, D) f32 [one row per E call], penalty (1, mel_codes) after the loop
    [, logits (n, mel_codes)])."""
    folded = [fold_layer(sd, i, cfg) for i in range(cfg.layers)]
    text = text_embed(text_ids, sd, cfg)
    first, gen_len = mel_embed([[cfg.start_mel]], 0, sd)
    hidden = torch.cat([torch.as_tensor(np.asarray(conds_latent)).float(), text, first], dim=1)        # graph D
    concat_len = hidden.shape[1]
    limit = cfg.max_generate - concat_len
    if max_new is not None:
        limit = min(limit, max_new)
    pen = torch.ones((1, cfg.mel_codes)) if penalty is None else torch.as_tensor(np.asarray(penalty)).float().clone().view(1, -1)
    H, hd = cfg.heads, cfg.head_dim
    pk = [torch.zeros((H, hd, 0)) for _ in range(cfg.layers)]
    pv = [torch.zeros((H, 0, hd)) for _ in range(cfg.layers)]
    ids_out, hid_out, log_out = [], [], []
    flag, reset, n = 1, 0, 0
    while n < limit:
        pk, pv, last, mid, logits = step_e(folded, pk, pv, pen, hidden, flag, sd, cfg)
        tok = int(mid.view(-1)[0])
        ids_out.append(tok)
        hid_out.append(last.view(-1).clone())
        log_out.append(logits.view(-1).clone())
        n += 1
        if tok == cfg.stop_mel:
            break
        flag = 0
        pen[:, tok] = cfg.repeat_penalty
        if n > cfg.penalty_range and ids_out[reset] != tok:
            pen[:, ids_out[reset]] = 1.0
            reset += 1
        hidden, gen_len = mel_embed([[tok]], gen_len, sd)
    out = (np.asarray(ids_out, dtype=np.int32), torch.stack(hid_out).numpy(), pen.numpy())
    return out + (torch.stack(log_out).numpy(),) if return_logits else out
