/*
 * dualar.h -- C-ABI of the B200-native dual-AR decode engine (libdualar.so).
 *
 * This is the drop-in boundary for ONE path of smolGura/fish-tts: the Fish-Speech
 * DualARTransformer autoregressive decode step and the token loop around it.  The reference has
 * no FFI of its own (it is pure Python); the seam it offers is the `decode_one_token` callable
 * returned by `init_model` (fish_tts/models/inference.py:387-414) and threaded through
 * `generate` / `generate_streaming` / `decode_n_tokens[_streaming]` (inference.py:158-384,
 * 643-738).  Each entry point below names the reference interface it stands in for.  Only plain
 * pointers and sizes cross this boundary -- no torch types.  INTEGRATION.md shows the ctypes
 * binding and the three-line patch-in for fish_tts/synthesizer.py.
 *
 * Conventions
 *   - every call returns 0 on success, a negative DUALAR_E* code otherwise; the message is
 *     available from dualar_last_error() (thread-local).  CUDA faults are reported by the first
 *     call that synchronises (the reference raises Python exceptions on the calling thread,
 *     SURVEY.md section 8b "Errors").
 *   - weights, activations and KV caches are bf16 (the FishTTS default precision,
 *     synthesizer.py:93,123-128); token ids are int32 (inference.py:27).
 *   - one engine = one request at a time, like one reference model instance (shared KV cache,
 *     synthesizer.py:44,699).  Different engines are independent and may live on different GPUs.
 *   - `stream` arguments are `cudaStream_t` passed as void*; NULL means the legacy default stream.
 */
#ifndef DUALAR_H_
#define DUALAR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DUALAR_ABI_VERSION 1

enum {
  DUALAR_OK = 0,
  DUALAR_EINVAL = -1,   /* bad argument / unsupported configuration           */
  DUALAR_ECUDA = -2,    /* a CUDA runtime call failed                          */
  DUALAR_ESTATE = -3,   /* call out of order (e.g. decode before prefill)      */
  DUALAR_EMISSING = -4, /* a weight the configuration needs was never loaded   */
  DUALAR_EDEVICE = -5   /* device-side fault flag raised by a kernel           */
};

/* Field-for-field the reference's DualARModelArgs (fish_tts/models/llama.py:31-123) plus the three
 * tokenizer-derived ids that enter the hot path (llama.py:418-420, inference.py:123,182). */
typedef struct dualar_config {
  int32_t abi_version; /* DUALAR_ABI_VERSION */
  int32_t vocab_size, n_layer, n_head, dim, intermediate_size, n_local_heads, head_dim;
  int32_t max_seq_len;
  int32_t codebook_size, num_codebooks;
  int32_t n_fast_layer, fast_dim, fast_n_head, fast_n_local_heads, fast_head_dim,
      fast_intermediate_size;
  int32_t tie_word_embeddings, attention_qkv_bias, attention_o_bias, attention_qk_norm;
  int32_t fast_attention_qkv_bias, fast_attention_o_bias, fast_attention_qk_norm;
  int32_t scale_codebook_embeddings;
  int32_t semantic_begin_id, semantic_end_id, im_end_id;
  float rope_base, norm_eps;
} dualar_config;

typedef struct dualar_engine dualar_engine;

/* ---- lifetime ------------------------------------------------------------------------------
 * Replaces: DualARTransformer.from_pretrained + model.to(device, dtype) (llama.py:466-500,
 * inference.py:394-395) and BaseTransformer/DualARTransformer.setup_caches (llama.py:378-398,
 * 544-559).  Weights are handed over under the reference's own state_dict keys
 * ("layers.3.attention.wqkv.weight", ...), as bf16, from host or device memory; the engine
 * repacks them once into its own HBM arena (w1/w3 row-interleaved, fast stack contiguous). */
int dualar_create(const dualar_config *cfg, int device, dualar_engine **out);
int dualar_load_weight(dualar_engine *e, const char *key, const void *data, int64_t n_elements,
                       int data_on_device);
/* checks completeness, builds the RoPE tables (llama.py:594-603), allocates KV caches unless bound
 * externally, captures the CUDA graphs. */
int dualar_finalize(dualar_engine *e);
void dualar_destroy(dualar_engine *e);
const char *dualar_last_error(void);

/* Optional, before dualar_finalize: run on KV buffers the caller owns, laid out exactly like the
 * reference's KVCache (llama.py:126-149): (1, n_local_heads, max_seq_len, head_dim) bf16 for slow
 * layers, (1, fast_n_local_heads, num_codebooks, fast_head_dim) for fast layers.  This is what lets
 * the step run beneath an unmodified reference prefill ("phase A"). */
int dualar_bind_kv(dualar_engine *e, int is_fast, int layer, void *k_cache, void *v_cache);

/* ---- the step: replaces decode_one_token_ar (inference.py:83-155) ----------------------------
 * All pointers are DEVICE pointers so nothing synchronises with the host, like the reference's
 * compiled step:
 *   x               (num_codebooks+1) int32   -- the current token column (`x[0, :, 0]`)
 *   input_pos       1 int32                   -- `input_pos[0]`
 *   previous_tokens (num_codebooks+1, 16) int32 view with row stride `prev_row_stride` elements
 *                   (the caller's window slice, inference.py:186-191); NULL = no repetition penalty
 *                   (the prefill call, inference.py:353-362)
 *   temperature, top_p, repetition_penalty    -- 1 float each (the reference's 0-dim tensors)
 *   noise           optional bf16 Exp(1) draws for `multinomial_sample_one_no_sync`
 *                   (inference.py:24-27): vocab_size values for the slow head followed by
 *                   (num_codebooks-1) * min(1024, codebook_size) for the fast heads; NULL = draw
 *                   them in-kernel from the engine's Philox stream (dualar_seed).
 *   out             (num_codebooks+1) int32   -- [semantic vocab id, cb0 ... cb_{C-1}]
 * The call is asynchronous on `stream`. */
int dualar_step(dualar_engine *e, const int32_t *x, const int32_t *input_pos,
                const int32_t *previous_tokens, int64_t prev_row_stride, const float *temperature,
                const float *top_p, const float *repetition_penalty, const void *noise,
                int32_t *out, void *stream);

/* ---- the loop: replaces generate / generate_streaming (inference.py:279-384, 643-738) --------
 * HOST buffers in, HOST buffers out.
 *   prompt  (num_codebooks+1, prompt_len) int32 row-major, as ContentSequence.encode_for_inference
 *           builds it (inference.py:611-640)
 * dualar_prefill copies the prompt to the device, fills the KV cache for positions
 * [0, prompt_len) and produces the first token (no repetition penalty, inference.py:353-362).
 * dualar_decode enqueues up to `n_steps` further decode steps without any host round trip; a
 * device-side flag stops the work after <|im_end|> (inference.py:210-211) or `max_new_tokens`.
 * dualar_collect waits for the enqueued work and copies out every token column produced since
 * the prefill: out is (num_codebooks+1, out_capacity) int32 row-major, *n_tokens columns are
 * valid, *finished tells whether the request has ended.  (The reference's batch `generate`
 * returns prompt + these columns, and its caller drops the last one, inference.py:839.) */
int dualar_prefill(dualar_engine *e, const int32_t *prompt, int prompt_len, int max_new_tokens,
                   float temperature, float top_p, float repetition_penalty, void *stream);
int dualar_decode(dualar_engine *e, int n_steps, void *stream);
int dualar_collect(dualar_engine *e, int32_t *out, int out_capacity, int *n_tokens, int *finished,
                   void *stream);
/* Streaming hand-off (generate_streaming, inference.py:643-738 / synthesize_stream, synthesizer.py:483-584): `n_steps` more
 * decode steps, then -- on the same stream, with no host wait in between -- an asynchronous copy of every column produced
 * and not yet handed out into the caller's (pinned) HOST buffer host_out, laid out (num_codebooks+1, n_steps+1) int32
 * row-major, and an event.  host_state is 5 int32 in (pinned) host memory: [0] columns generated so far, [1] request
 * finished, [2] device fault flag -- valid once dualar_wait reports the ticket -- and, set before the call returns, [3]
 * index of the first column of this chunk, [4] columns copied (some may lie beyond [0] after <|im_end|>).  The host can
 * enqueue the next chunk before waiting for this one, so the decode stream never idles while the codec consumes a chunk.
 * dualar_wait: block = 1 waits for the chunk, block = 0 polls (returns 1 while it has not landed). */
int dualar_decode_async(dualar_engine *e, int n_steps, int32_t *host_out, int32_t *host_state, void *stream, int *ticket);
int dualar_wait(dualar_engine *e, int ticket, int block);
/* prefill + decode to the end + collect in one call (what `generate` does). */
int dualar_generate(dualar_engine *e, const int32_t *prompt, int prompt_len, int max_new_tokens,
                    float temperature, float top_p, float repetition_penalty, int32_t *out,
                    int out_capacity, int *n_tokens, void *stream);

/* ---- batched decode: B requests advance together (BASELINE configs[3], "batched decode bs=32") -----------------
 * The reference is batch 1 only (inference.py:73 reads `logits[0, -1]`, :355 views the prompt as (1, C+1, T)), so there is
 * no reference entry point to mirror: these calls are the loop API above with a `slot` argument.  A slot is one request:
 * its own KV cache (`slot_seq_len` positions), position, repetition window, sampling parameters and Philox stream, so its
 * tokens are those of the same request run alone.  The linears are tcgen05 GEMMs (csrc/gemm_tc.cuh) with the requests as
 * the N dimension: one decode step of a GROUP of slots streams every weight byte once for the whole group.
 * Slots live in groups of `batch_group_slots` (dualar_set_option, default 32, <= 128): slot s belongs to group
 * s / batch_group_slots; every group has its own buffers, graph and stream and the groups of a step run concurrently (one
 * group's step is a latency-bound chain of ~540 kernels that leaves the GPU mostly idle).  Options (dualar_set_option):
 * "batch_group_slots" (before dualar_batch_init), "batch_persistent" (1: the step as two persistent cooperative launches,
 * csrc/bstep.cuh -- identical results, measured slower, an experiment switch), "batch_decode_join" (default 1; 0: dualar_batch_decode
 * does not make the caller's stream wait for the groups -- dualar_batch_read / dualar_batch_collect wait for what they read -- so a
 * serving loop can collect finished requests and enqueue prefills while the next burst of steps runs).
 *   dualar_batch_init     after dualar_finalize: allocates `max_batch` slots (<= 1024) and captures one CUDA graph per group
 *   dualar_batch_prefill  HOST prompt (num_codebooks+1, prompt_len) int32 -> KV rows of the slot through the tensor-core
 *                         prefill; the slot then joins the batch and the NEXT dualar_batch_decode step produces its first
 *                         token.  `noise`: optional explicit Exp(1) draws (device, layout of dualar_set_noise), NULL = Philox(seed).
 *                         With several groups the prefill runs asynchronously on the engine's own stream beside the decoding
 *                         groups and the call returns at once; with one group it is synchronous on `stream`
 *   dualar_batch_decode   n_steps batched steps of every group that holds a request, no host round trip; returns at once, the
 *                         caller's stream waits for all groups; finished or empty slots ride along as no-ops
 *   dualar_batch_collect  like dualar_collect, for one slot
 *   dualar_batch_release  frees the slot for the next request (continuous batching)
 *   dualar_batch_read     introspection (tests, bench): "slow_logits_raw" / "slow_logits" (B x vocab bf16), "hidden" (B x dim),
 *                         "fast_logits" (B x (num_codebooks-1) x fast_vocab), "tokens" (B x (num_codebooks+1) int32),
 *                         "positions" / "done" / "n_gen" (B int32) -- all in slot order over all groups; "launches" (1 int32:
 *                         kernels per step, summed over the groups), "groups" (1 int32) */
int dualar_batch_init(dualar_engine *e, int max_batch, int slot_seq_len);
int dualar_batch_prefill(dualar_engine *e, int slot, const int32_t *prompt, int prompt_len, int max_new_tokens,
                         float temperature, float top_p, float repetition_penalty, uint64_t seed, const void *noise,
                         void *stream);
int dualar_batch_decode(dualar_engine *e, int n_steps, void *stream);
int dualar_batch_collect(dualar_engine *e, int slot, int32_t *out, int out_capacity, int *n_tokens, int *finished,
                         void *stream);
int dualar_batch_release(dualar_engine *e, int slot);
int dualar_batch_read(dualar_engine *e, const char *name, void *host_dst, int64_t n_bytes, void *stream);

/* ---- sampling noise ---------------------------------------------------------------------------
 * The reference draws `q ~ Exp(1)` with torch's generator (inference.py:26).  Here the draws come
 * from a counter-based Philox4x32-10 stream keyed by (seed, step, head, element), so any process
 * can reproduce them.  dualar_fill_noise writes exactly the values the in-kernel path would use
 * for step `step` and head `head` (0 = slow head, k>=1 = fast head of codebook k) as bf16 into a
 * device buffer -- feed them to the reference/oracle for shared-RNG parity. */
int dualar_seed(dualar_engine *e, uint64_t seed);
int dualar_fill_noise(dualar_engine *e, uint64_t seed, uint32_t step, uint32_t head, void *out_bf16,
                      int64_t n, void *stream);
/* explicit per-step noise for the loop API: `noise` is a device buffer holding, for each step s in
 * order, the layout described at dualar_step; NULL switches back to the Philox stream. */
int dualar_set_noise(dualar_engine *e, const void *noise, int64_t n_steps);

/* ---- options ----------------------------------------------------------------------------------
 * "cpu_scalar_semantics" (0/1, default 0): torch evaluates `bf16_tensor op fp32_0dim_tensor`
 *     (logits / temperature, score * repetition_penalty; inference.py:42-44, 58) and
 *     `x / math.sqrt(C + 1)` (llama.py:427-429) differently on its two devices -- CUDA kernels cast the
 *     scalar to bf16 / multiply by the reciprocal, CPU kernels keep it in fp32 / divide.  0 follows the
 *     reference on CUDA (what a GPU user of the reference gets), 1 follows it on the CPU (what the
 *     golden fixtures were generated with).
 * "candidate_delta" (default 6.0, before finalize): slow-head candidates are logits >= max - delta.
 * "fast_qkv_table" (0/1, before finalize; default 1): passes >= 1 of the fast stack start from the embedding of a
 *     code, so the first fast layer's q|k|v is a function of that code alone; it is tabulated at finalize
 *     (codebook_size x fast qkv rows, bf16; 16 MB for s1-mini) and that wqkv phase is dropped from the step.
 * "prefill_mode" (0/1, default 0): 0 = dualar_prefill pushes prompt positions [0, T-1) through the tensor-core GEMMs in chunks of
 *     256 positions (the reference prefills in one forward, inference.py:353-362); 1 = one position per launch through the decode
 *     kernel (the round-1 path, kept as a cross-check).
 * "prefix_reuse" (0/1, default 1): dualar_prefill keeps the KV rows of the prompt positions shared with the previous request's
 *     prompt (the prefilled VoiceProfile references, synthesizer.py:363-377) and prefills only from the first differing position.
 * "mega_kernel" (0/1, before finalize; default 1): run the whole decode step as ONE persistent cooperative
 *     kernel (csrc/mega.cuh) instead of one kernel per phase (kept as a cross-check).  The two sum their dot
 *     products in different fp32 orders (tensor-core chunks vs. FMA chains), so logits agree to bf16 rounding,
 *     not bit for bit; the samplers of both are exact on the logits they are given. */
int dualar_set_option(dualar_engine *e, const char *name, double value);

/* ---- introspection (tests, bench) -------------------------------------------------------------
 * names: "slow_logits" (vocab_size bf16, after the repetition penalty), "slow_logits_raw"
 * (before it), "hidden" (dim bf16, the un-normalised last-layer output), "fast_logits"
 * ((num_codebooks-1) x fast_vocab bf16, before the penalty), "tokens" (num_codebooks+1 int32),
 * "nucleus" (num_codebooks int32: how many candidates survived top-p per head); and the per-layer
 * scratch of the LAST slow layer executed: "qkv", "y", "h", "act" (bf16), plus "fast_x" / "fast_in"; and two host-side
 * counters of the last dualar_prefill (1 int32 each): "prefix_reused" (prompt positions whose KV rows were kept),
 * "prefill_launches" (kernels the tensor-core prefill launched). */
int dualar_read_buffer(dualar_engine *e, const char *name, void *host_dst, int64_t n_bytes,
                       void *stream);
/* test hook: run the sampler (inference.py:30-80) of head `head` on caller-supplied bf16 logits
 * (vocab_size of them for head 0, min(1024, codebook_size) for a fast head).  Device pointers as in
 * dualar_step; writes the sampled index to out_token (1 int32, device).
 * With option mega_kernel = 1 (the default) this runs ONE WHOLE STEP of the persistent decode kernel whose logits epilogues
 * are fed the caller's logits, so the product path's penalty, statistics, candidate list and samplers are what is tested
 * (the engine's KV row at the current position is overwritten: prefill again before decoding); with mega_kernel = 0 it runs
 * the per-phase kernels' sampler. */
int dualar_debug_sample(dualar_engine *e, int head, const void *logits, const int32_t *previous_tokens,
                        int64_t prev_row_stride, const float *temperature, const float *top_p,
                        const float *repetition_penalty, const void *noise, int32_t *out_token,
                        void *stream);
/* kernels launched per decode step / per prefill position (what one graph replay contains) */
int dualar_launches_per_step(const dualar_engine *e, int *decode_step, int *prefill_position);
/* bytes of weights the engine holds, and how many of them belong to the fast stack */
int dualar_weight_bytes(const dualar_engine *e, int64_t *total, int64_t *fast);
int dualar_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DUALAR_H_ */
