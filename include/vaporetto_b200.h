/* vaporetto_b200 — C ABI of the B200-native `Predictor::predict` path.
 *
 * The reference (daac-tools/vaporetto, crate `vaporetto` 0.6.5) has no FFI layer: its surface is the
 * Rust API `Model` / `Predictor` / `Sentence` (vaporetto/src/lib.rs:82-91).  Every entry point below
 * names the Rust item it stands in for (paths relative to the reference's vaporetto/src/).  A Rust shim
 * that keeps the crate API and forwards to these symbols is sketched in INTEGRATION.md.
 *
 * Conventions
 *  - every function returns a vpt_status (0 = ok); the message of the last failure on the calling thread
 *    is available from vpt_last_error() (mirrors `VaporettoError`'s Display, errors.rs:41-56).
 *  - no panics/aborts cross the ABI; CUDA failures are reported as VPT_CUDA_ERROR.
 *  - a predictor is immutable after creation and may be shared by many host threads
 *    (`Predictor: Send + Sync`, shared as Arc<Predictor> in vaporetto_tantivy/src/lib.rs:62-67);
 *    each call uses its own CUDA stream and staging buffers.
 *  - there is no CPU fallback: without a usable CUDA device every compute entry point fails.
 */
#ifndef VAPORETTO_B200_H
#define VAPORETTO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef enum vpt_status {
    VPT_OK = 0,
    VPT_INVALID_MODEL = 1,    /* VaporettoError::InvalidModel    (errors.rs:16) */
    VPT_INVALID_ARGUMENT = 2, /* VaporettoError::InvalidArgument (errors.rs:19) */
    VPT_INVALID_SENTENCE = 3, /* VaporettoError::InvalidSentence (errors.rs:22) */
    VPT_DECODE_ERROR = 4,     /* VaporettoError::DecodeError     (errors.rs:34) */
    VPT_IO_ERROR = 5,         /* VaporettoError::IOError         (errors.rs:28) */
    VPT_CUDA_ERROR = 16,
    VPT_UNSUPPORTED = 17,
    VPT_INTERNAL = 18
} vpt_status;

/* per-sentence status written by the batch entry points (what `Sentence::from_raw` / `update_raw`
 * would have rejected, sentence.rs:174-189, plus malformed UTF-8 which a Rust &str cannot hold) */
enum {
    VPT_SENT_OK = 0,
    VPT_SENT_EMPTY = 1,        /* "text: must contain at least one character" */
    VPT_SENT_NUL = 2,          /* "text: must not contain NULL" */
    VPT_SENT_BAD_UTF8 = 3
};

#define VPT_NO_PATTERN 0xFFFFFFFFu /* u32::MAX in char_pma_states / type_pma_states (boundary_tag_scorer.rs:122-123) */

typedef struct vpt_model vpt_model;
typedef struct vpt_predictor vpt_predictor;

/* Message of the last error raised on this thread ("" if none). */
const char* vpt_last_error(void);

/* ---- Model --------------------------------------------------------------------------------------- */

/* `Model::read_slice(&[u8]) -> Result<(Model, &[u8])>` (model.rs:127-134) and `Model::read` (model.rs:142-153).
 * `data` is the raw (already un-zstd'd) model image; `*consumed` (nullable) receives the bytes used. */
int vpt_model_read(const uint8_t* data, size_t len, vpt_model** out, size_t* consumed);

/* The model loading of the reference's `predict` command (predict/src/main.rs:110-111:
 * `Model::read(&mut zstd::Decoder::new(File::open(path)?)?)`): `data` is the content of a *.model.zst file.  The zstd
 * frames are decoded inside the library (libzstd.so.1, opened at run time; IOError when it is absent); an image that
 * does not start with a zstd magic number is read as a raw model, as `vpt_model_read` does. */
int vpt_model_read_zstd(const uint8_t* data, size_t len, vpt_model** out);

/* `KyteaModel::read` + `Model::try_from(KyteaModel)` (kytea_model.rs:423-450, :453-550; the reference's
 * `convert_kytea_model` tool): a KyTea binary model becomes a vaporetto model — the word-segmentation linear model's
 * character / type n-gram weights (i16 -> i32, truncated to the window), bias, and the dictionary words with their
 * left / inside / right weights by word-length bucket; KyTea's tag models are not converted (as in the reference). */
int vpt_model_read_kytea(const uint8_t* data, size_t len, vpt_model** model_out);

/* `Model::dictionary` (model.rs:155-158) / `WordWeightRecord::get_word, get_weights, get_comment` (dict_model.rs:53-66):
 * record `index` of the model's word dictionary; the pointers stay valid until the model is freed or its dictionary
 * replaced. */
uint64_t vpt_model_dictionary_len(const vpt_model* model);
int vpt_model_dictionary_get(const vpt_model* model, uint64_t index, const char** word, const int32_t** weights,
                             uint64_t* n_weights, const char** comment);

/* `Model::replace_dictionary` (model.rs:160-163) with the check of `WordWeightRecord::new` (dict_model.rs:39-50):
 * every record needs chars(word) + 1 weights, else VPT_INVALID_ARGUMENT and the model is unchanged.  `comments` (and
 * its elements) may be NULL.  With vpt_model_to_vec this is the reference's `manipulate_model` tool. */
int vpt_model_replace_dictionary(vpt_model* model, const char* const* words, const int32_t* const* weights,
                                 const uint64_t* n_weights, const char* const* comments, uint64_t n_records);

/* `Model::to_vec` / `Model::write` (model.rs:99-120): the model file image (magic + bincode standard encoding);
 * byte-identical to what the reference writes for the same model.  Release the buffer with vpt_blob_free. */
int vpt_model_to_vec(const vpt_model* model, uint8_t** bytes_out, uint64_t* len_out);
void vpt_model_free(vpt_model* model);

/* ---- Predictor ----------------------------------------------------------------------------------- */

/* `Predictor::new(model: Model, predict_tags: bool) -> Result<Predictor>` (predictor.rs:450-508).
 * Consumes `model` (it is freed, success or failure), builds the merged weight rows and the flat device
 * tables, and uploads them to CUDA device `device`.  device = -1 creates a host-only handle (tag prediction and
 * Sentence helpers work, every scoring entry point fails with VPT_CUDA_ERROR — there is no CPU scoring path). */
int vpt_predictor_new(vpt_model* model, int predict_tags, int device, vpt_predictor** out);
void vpt_predictor_free(vpt_predictor* predictor);

typedef struct vpt_predictor_info {
    int32_t device;
    int32_t predict_tags;       /* created with predict_tags = true */
    int32_t n_tags;             /* Sentence::n_tags after fill_tags (predictor.rs:553) */
    int32_t char_scorer;        /* 0 none, 1 Boundary, 2 BoundaryTag       (char_scorer.rs:84-89) */
    int32_t type_scorer;        /* 0 none, 1 Boundary, 2 BoundaryCache, 3 BoundaryTag (type_scorer.rs:92-101) */
    int32_t fast_path;          /* 1: k_score_fast (all rows inline), 0: k_score_general */
    int32_t bias;
    int32_t char_window, type_window;
    uint32_t n_char_patterns, n_type_patterns;
    uint32_t n_char_nodes, n_type_nodes;
    uint32_t max_char_pattern_len;
    uint64_t blob_bytes;        /* size of the device-resident model */
    int32_t kernel_launches_per_batch;
} vpt_predictor_info;
int vpt_predictor_get_info(const vpt_predictor* predictor, vpt_predictor_info* out);

/* Flat device model: serialise on one rank, broadcast as bytes (NCCL), rebuild on the others.
 * (Replaces `Predictor::serialize_to_vec` / `deserialize_from_slice_unchecked`, predictor.rs:640-664, whose
 * daachorse-private layout is not reproducible; the blob format is this library's own.)
 * Predictors made from a blob score boundaries; tag prediction needs vpt_predictor_new. */
/* Host-only build of the flat model (no CUDA device needed): what vpt_predictor_new uploads.  Consumes `model`.
 * The returned buffer is released with vpt_blob_free. */
int vpt_blob_build(vpt_model* model, int predict_tags, uint8_t** blob_out, uint64_t* len_out);
void vpt_blob_free(uint8_t* blob);
uint64_t vpt_predictor_blob_size(const vpt_predictor* predictor);
int vpt_predictor_blob_export(const vpt_predictor* predictor, void* dst, uint64_t capacity);
int vpt_predictor_from_blob(const void* blob, uint64_t len, int device, vpt_predictor** out);

/* ---- predict ------------------------------------------------------------------------------------- */

/* Batched `Predictor::predict(&self, &mut Sentence)` (predictor.rs:518-543) over HOST buffers.
 *
 * Sentence i is utf8[byte_offsets[i] .. byte_offsets[i+1]) (raw text, as given to `Sentence::from_raw`).
 * With n_i = number of characters of sentence i, its boundaries occupy
 *   scores_out / boundaries_out [ bound_offsets_out[i] .. bound_offsets_out[i] + max(n_i - 1, 0) )
 * (`Sentence::boundary_scores()`, sentence.rs:1040-1046; `Sentence::boundaries()` as 0 = NotWordBoundary,
 * 1 = WordBoundary, sentence.rs:70-82) and, when requested, its pattern states occupy
 *   char_states_out / type_states_out [ char_offsets_out[i] .. + n_i )
 * (`char_pma_states` / `type_pma_states`, sentence.rs:92-93; only meaningful for a predictor created with
 * predict_tags = true on a model that has tag models).
 * Rejected sentences (status_out[i] != 0) keep their slots, filled with zeros / VPT_NO_PATTERN.
 *
 * out_capacity / states_capacity are the element capacities of the output arrays; if too small the call
 * fails with VPT_INVALID_ARGUMENT and the required sizes are in *n_boundaries_out / *n_chars_out.
 * scores_out, char_states_out, type_states_out, char_offsets_out, status_out may be NULL (without scores_out the
 * inline-row kernel skips the score stores and only boundaries cross PCIe).
 * For full PCIe bandwidth pass page-locked (pinned) host buffers.  The batch flows through an internal pipeline
 * (copy-in + kernels and copy-out on separate streams, four chunks in flight, chunk sizes ramping up from 1/8 of
 * env VPT_CHUNK_SENTENCES, default 262144); VPT_TRACE=1 prints its per-chunk timeline to stderr. */
int vpt_predict_batch(const vpt_predictor* predictor, const uint8_t* utf8, const uint64_t* byte_offsets, size_t n_sent,
                      int32_t* scores_out, uint8_t* boundaries_out, size_t out_capacity, uint64_t* bound_offsets_out,
                      int32_t* status_out, uint32_t* char_states_out, uint32_t* type_states_out,
                      size_t states_capacity, uint64_t* char_offsets_out, uint64_t* n_boundaries_out,
                      uint64_t* n_chars_out);

/* Same over DEVICE buffers, asynchronous on `cuda_stream` (a cudaStream_t; NULL = default stream).
 * d_utf8 must be 16-byte aligned and its allocation readable up to the next multiple of 16 bytes.
 * d_workspace: vpt_workspace_size(n_sent) bytes of scratch.  d_scores/d_boundaries must hold the batch's
 * total boundary count (an upper bound is total_bytes - 1); d_bound_offsets [n_sent+1];
 * d_status [n_sent]; d_char_states / d_type_states / d_char_offsets nullable. */
uint64_t vpt_workspace_size(size_t n_sent);

/* PCI bus id ("0000:1b:00.0") of CUDA device `device` as this process sees it (CUDA_VISIBLE_DEVICES applied): what a
 * launcher needs to bind a rank to the GPU's NUMA node (/sys/bus/pci/devices/<id>/numa_node).  buf: >= 16 bytes. */
int vpt_device_pci_bus_id(int device, char* buf, size_t capacity);
int vpt_predict_batch_dev(const vpt_predictor* predictor, const uint8_t* d_utf8, const uint64_t* d_byte_offsets,
                          size_t n_sent, void* d_workspace, uint64_t workspace_bytes, int32_t* d_scores,
                          uint8_t* d_boundaries, uint64_t* d_bound_offsets, int32_t* d_status,
                          uint32_t* d_char_states, uint32_t* d_type_states, uint64_t* d_char_offsets,
                          void* cuda_stream);

/* vpt_predict_batch_dev with CUDA events recorded on `cuda_stream` around each stage; synchronises the stream
 * and returns the device time of each stage in milliseconds: stage_ms[0] = k_count, [1] = k_scan_groups,
 * [2] = k_score_*  (bench.py's roofline leg times the dominant kernel with this). */
int vpt_predict_batch_dev_profiled(const vpt_predictor* predictor, const uint8_t* d_utf8,
                                   const uint64_t* d_byte_offsets, size_t n_sent, void* d_workspace,
                                   uint64_t workspace_bytes, int32_t* d_scores, uint8_t* d_boundaries,
                                   uint64_t* d_bound_offsets, int32_t* d_status, uint32_t* d_char_states,
                                   uint32_t* d_type_states, uint64_t* d_char_offsets, void* cuda_stream,
                                   float* stage_ms);

/* Single-sentence `Predictor::predict` (batch of one).  Returns VPT_INVALID_ARGUMENT with the reference's
 * message for an empty text or a text containing U+0000 (sentence.rs:174-189).  *n_chars_out receives n;
 * scores_out/boundaries_out need n-1 entries (capacity in elements), states n entries (nullable).
 * Sentences of up to 2 KiB on an inline-row model take a path without copy calls: the text goes into a pinned block the
 * kernel reads over PCIe, the results are stored into the same block: one launch + one synchronisation, about 26 us per call
 * on a B200 (a CPU scores such a sentence in a few us: batch the sentences, vpt_predict_batch / vpt_tokenize_lines). */
int vpt_predict(const vpt_predictor* predictor, const uint8_t* utf8, size_t n_bytes, int32_t* scores_out,
                uint8_t* boundaries_out, size_t out_capacity, uint32_t* char_states_out, uint32_t* type_states_out,
                size_t states_capacity, uint64_t* n_chars_out);

/* ---- tags on the device (`Predictor::predict_tags`, predictor.rs:546-637, for a whole batch) ------------------------
 * After vpt_predict_batch_dev with state outputs: for every character position i (indexed like the states, by
 * d_char_offsets) that ends a token known to the tag model, d_tag_token[i] = token id (else -1) and
 * d_tag_cand[i * n_tags + k] = index of the chosen candidate of tag slot k (else -1): the arrays vpt_fill_tags writes for
 * one sentence.  Token lookup (exact, by bytes), the weight vectors keyed by (pattern id, token, rel position) with the
 * reference's suffix merge (PositionalWeightWithTag +=, predictor.rs:242-262) and the per-slot arg-max (first strict
 * maximum, predictor.rs:286-304) run in one kernel, k_tags.  *d_unserved (nullable, zero it first) counts tokens whose
 * tag model exceeds the limits of the device tables (more than 64 scores or 8 tag slots; none of the reference's
 * models): their entries are -1 and vpt_fill_tags serves them.  Returns VPT_UNSUPPORTED when the whole model is beyond
 * the limits, VPT_INVALID_ARGUMENT for a predictor created with predict_tags = false (the reference panics). */
int vpt_predict_tags_batch_dev(const vpt_predictor* predictor, const uint8_t* d_utf8, const uint64_t* d_byte_offsets,
                               size_t n_sent, const int32_t* d_status, const uint8_t* d_boundaries,
                               const uint64_t* d_bound_offsets, const uint64_t* d_char_offsets,
                               const uint32_t* d_char_states, const uint32_t* d_type_states, int32_t* d_tag_token,
                               int32_t* d_tag_cand, uint32_t* d_unserved, void* cuda_stream);

/* Host-buffer form: predict + predict_tags for a batch; the pattern-id states never leave the device.  Outputs as
 * vpt_predict_batch (scores_out nullable) plus tag_token_out [chars], tag_cand_out [chars * n_tags] and
 * char_offsets_out [n_sent + 1]; chars_capacity in characters. */
int vpt_predict_batch_tags(const vpt_predictor* predictor, const uint8_t* utf8, const uint64_t* byte_offsets, size_t n_sent,
                           int32_t* scores_out, uint8_t* boundaries_out, size_t out_capacity, uint64_t* bound_offsets_out,
                           int32_t* status_out, int32_t* tag_token_out, int32_t* tag_cand_out, size_t chars_capacity,
                           uint64_t* char_offsets_out, uint64_t* n_boundaries_out, uint64_t* n_chars_out,
                           uint64_t* n_unserved_out);

/* ---- tags (host side; `Sentence::fill_tags` -> `Predictor::predict_tags`, predictor.rs:546-637) ------ */

/* For one sentence with final boundaries (0 not / 1 boundary / 2 unknown, sentence.rs:70-82) and the
 * states produced by predict: for every character position i that ends a token known to the tag model,
 * tag_token_out[i] = token id (else -1) and tag_cand_out[i*n_tags + k] = index of the chosen candidate of
 * tag slot k (else -1).  tag_scores_out (nullable, n_chars * score_stride) receives the raw score vectors
 * (`Predictor::store_tag_scores`, predictor.rs:512).  Fails with VPT_INVALID_ARGUMENT
 * ("this predictor is created with predict_tags = false") where the reference panics (predictor.rs:547-551). */
int vpt_fill_tags(const vpt_predictor* predictor, const uint8_t* utf8, size_t n_bytes, const uint8_t* boundaries,
                  const uint32_t* char_states, const uint32_t* type_states, int32_t* tag_token_out,
                  int32_t* tag_cand_out, int32_t* tag_scores_out, size_t score_stride);
/* tag string of (token id, slot, candidate); NULL if out of range. Valid for the predictor's lifetime. */
const char* vpt_tag_string(const vpt_predictor* predictor, uint32_t token_id, uint32_t slot, uint32_t cand);
/* number of candidates of a slot (0 if out of range) and the score-vector length of a token */
uint32_t vpt_tag_n_candidates(const vpt_predictor* predictor, uint32_t token_id, uint32_t slot);
uint32_t vpt_tag_score_len(const vpt_predictor* predictor, uint32_t token_id);
uint32_t vpt_tag_n_tokens(const vpt_predictor* predictor);

/* predict (+ fill_tags) of a batch with COMPACT results, for callers that want the segmentation and the tags rather
 * than the score strip: what crosses PCIe is one bit per boundary and one small record per token.
 *   boundary_bits_out  the boundaries of all sentences as one bit stream (bit k of the stream = bit k % 32 of word
 *                      k / 32; 1 = WordBoundary): sentence s owns the bits [B_s, B_s + max(n_chars_out[s], 1) - 1)
 *                      with B_s = the sum over the earlier sentences (CharacterBoundary values of
 *                      `Sentence::boundaries()`, sentence.rs:1016-1046; vpt_unpack_boundaries gives the byte form);
 *   n_chars_out[s]     characters of the sentence (a rejected sentence, status_out[s] >= 2, keeps its count and owns
 *                      zero bits; an empty one has 0); status_out[s] as in vpt_predict_batch;
 *   n_tokens_out[s]    (nullable unless tags are requested) tokens of the sentence = boundaries set + 1;
 *   token_ids_out / token_cands_out (both NULL: no tag prediction; needs predict_tags = true otherwise): token r of
 *                      sentence s in text order is record T_s + r (T_s = the sum of n_tokens_out over the earlier
 *                      sentences): the token id for vpt_tag_string (-1: the token has no tag model) and, per tag slot,
 *                      the chosen candidate as one byte (255: none) -- `Predictor::predict_tags`, predictor.rs:546-637,
 *                      the same choice vpt_predict_batch_tags reports per character.
 * The totals come back in *n_boundaries_out / *n_tokens_total_out; *n_unserved_out counts tokens whose tag model
 * exceeds the device limits (0 for the reference's models; vpt_fill_tags serves those).  Too small capacities return
 * InvalidArgument with the totals set. */
int vpt_predict_batch_compact(const vpt_predictor* predictor, const uint8_t* utf8, const uint64_t* byte_offsets,
                              size_t n_sent, uint32_t* boundary_bits_out, size_t bits_capacity_words,
                              uint32_t* n_chars_out, uint8_t* status_out, uint32_t* n_tokens_out, int32_t* token_ids_out,
                              uint8_t* token_cands_out, size_t token_capacity, uint64_t* n_boundaries_out,
                              uint64_t* n_tokens_total_out, uint64_t* n_unserved_out);
/* bits [first_bit, first_bit + n) of a boundary bit stream as bytes (0 / 1) */
int vpt_unpack_boundaries(const uint32_t* boundary_bits, uint64_t first_bit, uint64_t n, uint8_t* boundaries_out);

/* ---- Sentence helpers (host side) ----------------------------------------------------------------- */

/* `CharacterType::get_type` per character (sentence.rs:50-67) / `Sentence::char_types()` (sentence.rs:993).
 * Returns the reference's InvalidArgument errors for empty text / NUL.  *n_chars_out receives n. */
int vpt_char_types(const uint8_t* utf8, size_t n_bytes, uint8_t* types_out, size_t capacity, uint64_t* n_chars_out);

/* `SplitLinebreaksFilter::filter(&mut Sentence)` (vaporetto_rules/src/sentence_filters/split_linebreaks.rs:9-37) on the
 * host, for the Sentence API: the boundary on either side of every '\r' / '\n' becomes WordBoundary (1).  (The lines
 * path never sees these characters inside a sentence: it splits at them.) */
int vpt_split_linebreaks(const uint8_t* utf8, size_t n_bytes, uint8_t* boundaries, size_t n_boundaries);

/* `ConcatGraphemeClustersFilter::filter(&mut Sentence)` (vaporetto_rules/src/sentence_filters/
 * concat_grapheme_clusters.rs:10-35) on the host, for the Sentence API: `boundaries` (n_chars - 1 values, 0 / 1, as
 * vpt_predict returns them) loses every boundary inside an extended grapheme cluster of the text (UAX #29 as in
 * unicode-segmentation 1.12 `graphemes(true)`; the rule engine vpt_tokenize_lines runs on the device for
 * VPT_WSCONST_GRAPHEME). */
int vpt_concat_grapheme_clusters(const uint8_t* utf8, size_t n_bytes, uint8_t* boundaries, size_t n_boundaries);

/* `Sentence::write_tokenized_text` (sentence.rs:850-886): tokens joined by ' ', with '/tag' suffixes when
 * tag_token/tag_cand are given (NULL otherwise), escaping ' ', '\\', '/'.  Tokens adjacent to an Unknown
 * boundary are skipped.  Returns the byte length needed in *len_out; writes at most `capacity` bytes. */
int vpt_write_tokenized_text(const vpt_predictor* predictor, const uint8_t* utf8, size_t n_bytes,
                             const uint8_t* boundaries, const int32_t* tag_token, const int32_t* tag_cand,
                             char* buf, size_t capacity, uint64_t* len_out);

/* ---- Whole-buffer tokenisation: the reference CLI's loop on the device ----------------------------- */

/* The loop of the reference's `predict` CLI (predict/src/main.rs:126-181) over a whole buffer of raw file bytes:
 *   for line in stdin.lines():
 *       s = KyteaFullwidthFilter(line) unless --no-norm          (main.rs:98,154; vaporetto_rules kytea_fullwidth.rs)
 *       if s.update_raw(..).is_ok() { predict(s); copy the boundaries to the original line; write_tokenized_text }
 *       write "\n"
 * Lines are split ON THE DEVICE with `BufRead::lines` semantics ('\n' or "\r\n" terminated; the last line may
 * be unterminated; a trailing '\n' adds no empty line), scored by the same kernels as vpt_predict_batch (with the
 * full-width character map applied to the code points the kernels look up when no_norm == 0; the map is one
 * character to one character, so the boundaries apply to the original text), and the tokenised ORIGINAL text
 * (' ' between tokens; '\\' before ' ', '\\', '/': sentence.rs:850-886) is materialised on the device: the only
 * transfers are the input bytes in and the output bytes out.  Lines that update_raw rejects (empty, or
 * containing U+0000) produce an empty line as in the CLI; so do lines that are not valid UTF-8 (the CLI stops
 * with an I/O error on those).  Score printing (--scores, --tag-scores) is not part of this path; tags: below.
 * `no_norm`: the CLI flag of the same name (0 = apply KyteaFullwidthFilter, the CLI default).
 * `wsconst_types`: the CLI's `--wsconst D/R/H/T/K/O` options as a bit set, bit t for CharacterType t (VPT_WSCONST_*):
 * `KyteaWsConstFilter` (vaporetto_rules/src/sentence_filters/kytea_wsconst.rs:27-44) clears the boundary between two
 * characters of such a type after prediction (types of the filtered text when no_norm == 0, main.rs:157);
 * VPT_WSCONST_GRAPHEME is `--wsconst G`: `ConcatGraphemeClustersFilter` (sentence_filters/concat_grapheme_clusters.rs:
 * 10-35) clears the boundaries inside every extended grapheme cluster (UAX #29 as in unicode-segmentation 1.12).
 * `out` receives the output lines, each terminated by '\n' (at most 3 * n_bytes + n_lines bytes); *out_len
 * the number of bytes produced (also when `out_capacity` was too small, which returns InvalidArgument);
 * *n_lines the number of input lines.  Chunk size of the internal pipeline: env VPT_CHUNK_BYTES (16 MiB, with
 * smaller chunks at both ends); VPT_TRACE=1 prints the pipeline's per-chunk timeline to stderr. */
#define VPT_WSCONST_DIGIT (1u << 1)    /* --wsconst D */
#define VPT_WSCONST_ROMAN (1u << 2)    /* --wsconst R */
#define VPT_WSCONST_HIRAGANA (1u << 3) /* --wsconst H */
#define VPT_WSCONST_KATAKANA (1u << 4) /* --wsconst T */
#define VPT_WSCONST_KANJI (1u << 5)    /* --wsconst K */
#define VPT_WSCONST_OTHER (1u << 6)    /* --wsconst O */
#define VPT_WSCONST_GRAPHEME (1u << 7) /* --wsconst G: ConcatGraphemeClustersFilter */
int vpt_tokenize_lines(const vpt_predictor* predictor, const uint8_t* utf8, size_t n_bytes, int no_norm,
                       uint32_t wsconst_types, uint8_t* out, size_t out_capacity, uint64_t* out_len, uint64_t* n_lines);

/* The same loop with the CLI's `--predict-tags` (predict/src/main.rs:130-136,159-166): `fill_tags` on the sentence that
 * was predicted (after the post-filters), tags copied to the original line, `write_tokenized_text` with tags
 * (sentence.rs:850-886: every token is followed by '/' + tag for its tag slots up to the last one that has a tag, the
 * tag strings escaped like the surface).  Tag prediction (token lookup by the bytes of the pre-filtered token, tag
 * weights, arg-max) and the output with its tag strings run on the device; the predictor must have been created with
 * predict_tags = 1.  `out` needs room for the tags: at most 3 * n_bytes + n_lines + n_bytes * (longest tag suffix). */
int vpt_tokenize_lines_tags(const vpt_predictor* predictor, const uint8_t* utf8, size_t n_bytes, int no_norm,
                            uint32_t wsconst_types, uint8_t* out, size_t out_capacity, uint64_t* out_len, uint64_t* n_lines);

/* `KyteaFullwidthFilter` for one character (vaporetto_rules/src/string_filters/kytea_fullwidth.rs:13-118): the
 * same function the kernels apply (csrc/textnorm.hpp). */
uint32_t vpt_kytea_fullwidth(uint32_t code_point);

/* library build info, e.g. "vaporetto_b200 0.1.0 sm_100a" */
const char* vpt_version(void);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* VAPORETTO_B200_H */
