// IndexTTS GPT-2 acoustic model: graphs B-E of the reference (IndexTTS/Export_IndexTTS.py:203-289) and the greedy decode
// loop around them (IndexTTS/Inference_IndexTTS_ONNX.py:726-781), with the KV cache, the repeat-penalty vector and the
// loop state resident on the device. See gpt2.cu.
#pragma once
#include <cstdint>

#include "engine.cuh"

namespace b200tts {

struct GptModel;

// Build device layouts from engine.weights["igpt.*"] (index-tts / Hugging Face GPT2 state-dict names).
GptModel* gpt_build(Engine& e);
void gpt_free(GptModel* m);
int gpt_dim(const GptModel& m);
int gpt_layers(const GptModel& m);
int gpt_heads(const GptModel& m);
int gpt_mel_codes(const GptModel& m);
int gpt_max_rows(const GptModel& m);          // KV-cache capacity in rows (MAX_GENERATE_LENGTH)

// Graph B: [start, ids, stop] -> text_embedding + text_pos_embedding[:n+2]; d_out (n_text + 2, D) fp32 device.
void gpt_text_embed(Engine& e, const int* d_text_ids, int n_text, float* d_out);
// Graph C: mel_embedding(id) + mel_pos_embedding[gen_len]; d_out (1, D).
void gpt_mel_embed(Engine& e, const int* d_id, int gen_len, float* d_out);

// Graph E, one call: `rows` new rows of hidden state (d_hidden, (rows, D) fp32 device) appended after `history` cached rows
// (history must equal the resident length, or 0 to start a new sentence). mask_flag 1 = causal among the new rows (the
// reference's int8 table x flag; -128 additive, Export_IndexTTS.py:245,268). d_penalty: (mel_codes) fp32 device.
// -> d_last_hidden (D) = ln_f(last row), d_max_id (1) int32 = argmax(mel_head(final_norm(.)) * penalty).
void gpt_step(Engine& e, const float* d_hidden, int rows, int history, int mask_flag, const float* d_penalty, int precision,
              float* d_last_hidden, int* d_max_id);
// Cached keys / values of one layer in the reference's layouts: key (H, 64, S), value (H, S, 64), S = resident rows.
void gpt_kv_export(Engine& e, int layer, float* d_key, float* d_value);
int gpt_resident_rows(const GptModel& m);

// One sentence, whole loop on the device: conds_latent (cond_rows, D) and text ids -> up to max_new greedy mel tokens and
// the ln_f hidden state of every call. d_penalty (mel_codes, in/out) carries the penalty vector across sentences as the
// reference does. Returns the number of tokens produced (the stop token, if hit, is the last one).
int gpt_generate(Engine& e, const float* d_conds, int cond_rows, const int* d_text_ids, int n_text, int max_new, int precision,
                 float* d_penalty, int* d_ids_out, float* d_hidden_out);

}  // namespace b200tts
