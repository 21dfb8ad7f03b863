/* CPU restatement, in plain C, of the byte / integer rules of the scoring path -- TEST INFRASTRUCTURE ONLY (the checker of
 * plantcaduceus_b200/genome_io.py, tokenizer.py and of the device kernels tokenize_kernel / extract_windows_kernel /
 * embed_kernel's index math; nothing under plantcaduceus_b200/ may load it).  Built by __graft_entry__.build() with gcc into
 * oracle/_build/libhost_rules.so.  Pinned by outputs of the reference itself: tests/golden/reference_run/small_windows.json and
 * vcf_windows.npz are what the reference's seq_from_vcf returned for those inputs; tests/golden/example_ids.npz the ids of its
 * example table (tests/test_oracle.py).
 *
 * Each function cites the reference lines it follows (paths under /root/reference).
 */
#include <stdint.h>
#include <string.h>

static uint8_t upper_ascii(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }

/* src/zero_shot_score.py:49-62 (SequenceDataset.__getitem__): ids = tokenizer(sequence) -- one id per character through the
 * tokenizer's table -- then ids[tokenIdx] = mask_token_id.  `n` characters form windows of length L; mask_pos < 0: no mask. */
void pcad_oracle_tokenize(const uint8_t* ascii, int64_t n, const uint8_t* lut, int L, int mask_pos, int mask_id, uint8_t* ids) {
  for (int64_t i = 0; i < n; ++i) ids[i] = lut[ascii[i]];
  if (mask_pos >= 0 && L > 0)
    for (int64_t w = 0; w * L + mask_pos < n; ++w) ids[w * L + mask_pos] = (uint8_t)mask_id;
}

/* src/zero_shot_score.py:185-198 (seq_from_vcf), for one record at 0-based position pos0:
 *     addIdx = 512 - tokenIdx
 *     if pos - tokenIdx < 0:  seq = chrom[0 : pos + addIdx].upper().rjust(512, "N")
 *     else:                   seq = chrom[pos - tokenIdx : pos + addIdx].upper().ljust(512, "N")
 * with Python's slice clipping at the end of the chromosome; L stands for the 512.  Returns the number of bases the slice
 * held (the rest of out[0..L) is 'N'); a slice longer than L cannot occur for 0 <= tokenIdx < L. */
int pcad_oracle_extract_window(const uint8_t* chrom, int64_t chrom_len, int64_t pos0, int token_idx, int L, uint8_t* out) {
  const int64_t add = (int64_t)L - token_idx;
  int64_t lo, hi;
  int right_justify;
  if (pos0 - token_idx < 0) { lo = 0; hi = pos0 + add; right_justify = 1; }
  else { lo = pos0 - token_idx; hi = pos0 + add; right_justify = 0; }
  if (hi < 0) hi = 0;                 /* chrom[0:negative] would count from the end in Python; pos0 >= 0 keeps hi >= 1 */
  if (hi > chrom_len) hi = chrom_len;
  if (lo > hi) lo = hi;
  const int64_t n = hi - lo;
  memset(out, 'N', (size_t)L);
  uint8_t* dst = right_justify ? out + (L - n) : out;
  for (int64_t k = 0; k < n; ++k) dst[k] = upper_ascii(chrom[lo + k]);
  return (int)n;
}

/* The reverse-complement strand the RC half of the model sees ([EXT] RCPSEmbedding.forward: complement_map[flip(ids)];
 * complement map from pretrain/llmlib/architectures/models/mamba/caduceus.py:100-105): out[t] = comp[ids[L-1-t]]. */
void pcad_oracle_rc_ids(const uint8_t* ids, int L, const uint8_t* comp, uint8_t* out) {
  for (int t = 0; t < L; ++t) out[t] = comp[ids[L - 1 - t]];
}

/* src/format_VCF.sh:42-44 (awk | bedtools slop -l 255 -r 256): the 0-based half-open interval of a 1-based position,
 * clipped to the chromosome. */
void pcad_oracle_slop(int64_t pos1, int64_t chrom_len, int left, int right, int64_t* start, int64_t* end) {
  int64_t s = pos1 - 1 - left, e = pos1 + right;
  *start = s < 0 ? 0 : s;
  *end = e > chrom_len ? chrom_len : e;
}
