// gzip on the GPU for the resampler's output.
//
// The reference writes its resampled list through libmcpl and closes it with
// mcpl_closeandgzip_outfile (clib/src/resamplemcpl.c:30-46): the product is always a
// gzip file.  At 1.6e10 particles/s the records exist in HBM long before a host could
// deflate them (zlib level 1 on 16 cores: ~1e7 particles/s), so the compression runs
// on the device too: the record stream is cut into members of `member_bytes`, each
// member becomes one complete gzip member (RFC 1952) holding a single dynamic-Huffman
// deflate block (RFC 1951): Huffman-coded literals plus, for fixed-size records, one
// kind of LZ77 match -- "the last `tail_bytes` of this record repeat the previous
// record's" (distance = record size), which is what the constant time / weight /
// pdgcode fields of a resampled MCPL list are.  That brings the file size to within
// a few percent of level-1 zlib.  The members are laid out back to back.  Concatenated gzip members are a valid
// gzip file for zlib's gzread (libmcpl), gzip(1) and Python's gzip module.
//
//   pass 1 (gz_size_kernel)   per member: code-length sum per thread, CRC-32
//   pass 2 (gz_scan_kernel)   member sizes -> byte offsets (all multiples of 4: the
//                             gzip FNAME field carries 0-3 filler characters)
//   pass 3 (gz_encode_kernel) header, bit-packed codes, end-of-block, CRC-32, ISIZE
//
// Every member announces its own compressed size in a gzip FEXTRA subfield ("KD", 4 bytes,
// as BGZF does with "BC"), so a reader can hop from header to header and inflate the
// members in parallel (kdsource_b200/mcpl.py); plain gzip readers skip the field.
//
// One Huffman table serves every member: it is built on the host
// (kdsb200_gz_build_table) from a byte histogram of the first records
// (kdsb200_gz_histogram), with every symbol kept encodable.
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "kdsb200.h"

namespace kdsb {

constexpr int GZ_THREADS = 256;
constexpr uint32_t CRC_POLY = 0xEDB88320u;  // reflected CRC-32 (IEEE 802.3), as gzip uses

// device image of the table (uploaded at the start of every compress call)
struct GzDev {
  uint32_t entry[257];   // len << 16 | bit-reversed code
  uint32_t rec_bytes, tail_bytes;       // record-aware mode (0: plain byte stream)
  uint32_t match_bits, match_nbits;     // length code + extra + distance code + extra
  uint32_t hdr_bits;
  uint32_t hdr[64];
  uint32_t crc_tab[256];
  uint32_t crc_mat[6][32];  // advance-by-(chunk << j)-zero-bytes operators, j = 0..5
  uint32_t init_full, init_last;  // M_n(0xffffffff) for a full / the last member
};

// ---- host: Huffman code lengths, canonical codes, deflate block header --------
static void huffman_lengths(const uint64_t* freq, int n, int maxlen, uint8_t* len) {
  std::vector<uint64_t> f(freq, freq + n);
  for (;;) {
    // plain O(n^2) two-smallest merging; n <= 257
    struct Node { uint64_t w; int l, r; };
    std::vector<Node> nodes;
    std::vector<int> live;
    for (int i = 0; i < n; i++)
      if (f[i] > 0) {
        nodes.push_back({f[i], -1 - i, 0});
        live.push_back((int)nodes.size() - 1);
      }
    for (int i = 0; i < n; i++) len[i] = 0;
    if (live.size() == 1) {
      len[-1 - nodes[live[0]].l] = 1;
      return;
    }
    while (live.size() > 1) {
      auto by_w = [&](int a, int b) { return nodes[a].w < nodes[b].w || (nodes[a].w == nodes[b].w && a < b); };
      std::sort(live.begin(), live.end(), by_w);
      const int a = live[0], b = live[1];
      nodes.push_back({nodes[a].w + nodes[b].w, a, b});
      live.erase(live.begin(), live.begin() + 2);
      live.push_back((int)nodes.size() - 1);
    }
    // depth of every leaf
    std::vector<std::pair<int, int>> stack{{live[0], 0}};
    int deepest = 0;
    while (!stack.empty()) {
      auto [id, d] = stack.back();
      stack.pop_back();
      if (nodes[id].l < 0) {
        len[-1 - nodes[id].l] = (uint8_t)d;
        deepest = std::max(deepest, d);
      } else {
        stack.push_back({nodes[id].l, d + 1});
        stack.push_back({nodes[id].r, d + 1});
      }
    }
    if (deepest <= maxlen) return;
    for (int i = 0; i < n; i++)  // flatten the distribution and retry
      if (f[i] > 0) f[i] = (f[i] + 1) / 2;
  }
}

static void canonical_codes(const uint8_t* len, int n, uint32_t* code_rev) {
  uint32_t bl_count[16] = {0}, next[16] = {0};
  for (int i = 0; i < n; i++) bl_count[len[i]]++;
  bl_count[0] = 0;
  uint32_t c = 0;
  for (int b = 1; b < 16; b++) {
    c = (c + bl_count[b - 1]) << 1;
    next[b] = c;
  }
  for (int i = 0; i < n; i++) {
    code_rev[i] = 0;
    if (!len[i]) continue;
    uint32_t v = next[len[i]]++, r = 0;
    for (int b = 0; b < len[i]; b++) r |= ((v >> b) & 1u) << (len[i] - 1 - b);  // MSB-first -> LSB-first
    code_rev[i] = r;
  }
}

struct BitWriter {
  uint32_t* w;
  uint32_t nbits = 0, cap;
  BitWriter(uint32_t* words, uint32_t cap_bits) : w(words), cap(cap_bits) {}
  bool put(uint32_t v, int n) {
    for (int b = 0; b < n; b++, nbits++) {
      if (nbits >= cap) return false;
      if ((v >> b) & 1u) w[nbits >> 5] |= 1u << (nbits & 31);
    }
    return true;
  }
};

// GF(2) operators on the CRC register: column c of M = image of bit c
static void mat_square(const uint32_t* m, uint32_t* out) {
  for (int c = 0; c < 32; c++) {
    uint32_t v = m[c], r = 0;
    for (int b = 0; v; b++, v >>= 1)
      if (v & 1u) r ^= m[b];
    out[c] = r;
  }
}
static uint32_t mat_apply(const uint32_t* m, uint32_t v) {
  uint32_t r = 0;
  for (int b = 0; v; b++, v >>= 1)
    if (v & 1u) r ^= m[b];
  return r;
}
// operator "append n zero bytes" by square-and-multiply on the one-zero-bit operator
static void crc_advance_matrix(uint64_t nbytes, uint32_t* out) {
  uint32_t sq[32], acc[32], tmp[32];
  sq[0] = CRC_POLY;  // one zero bit: r -> (r >> 1) ^ (r & 1 ? poly : 0)
  for (int c = 1; c < 32; c++) sq[c] = 1u << (c - 1);
  for (int c = 0; c < 32; c++) acc[c] = 1u << c;  // identity
  uint64_t nb = nbytes * 8;
  while (nb) {
    if (nb & 1) {
      for (int c = 0; c < 32; c++) tmp[c] = mat_apply(sq, acc[c]);
      memcpy(acc, tmp, sizeof(acc));
    }
    mat_square(sq, tmp);
    memcpy(sq, tmp, sizeof(sq));
    nb >>= 1;
  }
  memcpy(out, acc, sizeof(acc));
}

// ---- device ---------------------------------------------------------------------
// does the tail of the record at member offset o repeat the previous record's tail?
// (the first record of a member has no history: every member is its own stream)
__device__ __forceinline__ bool gz_tail_match(const unsigned char* member, uint32_t o, uint32_t R,
                                              uint32_t S) {
  if (o < R) return false;
  const uint32_t* a = reinterpret_cast<const uint32_t*>(member + o + R - S);
  const uint32_t* b = reinterpret_cast<const uint32_t*>(member + o - S);
  bool eq = true;
  for (uint32_t k = 0; k < S / 4; k++) eq = eq && (__ldg(a + k) == __ldg(b + k));
  return eq;
}

// histogram of the symbols the encoder will emit: bytes 0..255, [257] = tail matches
__global__ void gz_hist_kernel(const unsigned char* __restrict__ data, uint64_t nbytes,
                               uint32_t member_bytes, uint32_t R, uint32_t S,
                               unsigned long long* __restrict__ hist) {
  __shared__ unsigned int h[258];
  for (int i = threadIdx.x; i < 258; i += blockDim.x) h[i] = 0;
  __syncthreads();
  const uint64_t step = R ? R : 4;
  for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * step; i + step <= nbytes;
       i += (uint64_t)gridDim.x * blockDim.x * step) {
    uint32_t lit = (uint32_t)step;
    if (R && S) {
      const uint64_t mbase = i / member_bytes * member_bytes;
      if (gz_tail_match(data + mbase, (uint32_t)(i - mbase), R, S)) {
        lit -= S;
        atomicAdd(&h[257], 1u);
      }
    }
    for (uint32_t o = 0; o < lit; o += 4) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(data + i + o);
      atomicAdd(&h[w & 255u], 1u);
      atomicAdd(&h[(w >> 8) & 255u], 1u);
      atomicAdd(&h[(w >> 16) & 255u], 1u);
      atomicAdd(&h[w >> 24], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 258; i += blockDim.x)
    if (h[i]) atomicAdd(&hist[i], (unsigned long long)h[i]);
}

// ---- the member's bytes in shared memory --------------------------------------------
// Each thread walks its own chunk of the member word by word (the bit offsets and the CRC
// are sequential inside a chunk), which from global memory is a 32-way scattered 4-byte
// load per step.  A full member is therefore copied into shared memory first, with
// coalesced loads, chunk c at word offset c * (CW + pad): pad makes the stride odd, so the
// 32 lanes of a warp -- all at the same position inside their chunks -- hit 32 banks.
constexpr uint32_t GZ_STAGE_MAX = 96u << 10;  // members up to this size are staged

struct GzSrc {
  const uint32_t* g;   // member in global memory
  const uint32_t* sm;  // staged copy (nullptr: read global)
  uint32_t inv, pad;   // ceil(2^32 / CW), 0 | 1
  __device__ __forceinline__ uint32_t word(uint32_t o) const {  // o = byte offset, % 4 == 0
    const uint32_t i = o >> 2;
    return sm ? sm[i + __umulhi(i, inv) * pad] : __ldg(g + i);
  }
};

__device__ __forceinline__ uint32_t gz_stage_words(uint32_t member_bytes) {
  const uint32_t cw = member_bytes / (4u * GZ_THREADS);
  return GZ_THREADS * (cw + ((cw & 1u) ? 0u : 1u));
}

// nb == member_bytes (a full member) and member_bytes <= GZ_STAGE_MAX
__device__ __forceinline__ GzSrc gz_stage(const unsigned char* member, uint32_t member_bytes,
                                          uint32_t* sm) {
  const uint32_t cw = member_bytes / (4u * GZ_THREADS), pad = (cw & 1u) ? 0u : 1u;
  const uint32_t inv = (uint32_t)((0x100000000ull + cw - 1) / cw);
  const uint32_t* g = reinterpret_cast<const uint32_t*>(member);
  for (uint32_t i = threadIdx.x; i < cw * GZ_THREADS; i += GZ_THREADS)
    sm[i + __umulhi(i, inv) * pad] = __ldg(g + i);
  __syncthreads();
  return GzSrc{g, sm, inv, pad};
}

// does the tail of the record at member offset o repeat the previous record's tail?
// (the first record of a member has no history: every member is its own stream)
__device__ __forceinline__ bool gz_tail_match(const GzSrc& src, uint32_t o, uint32_t R, uint32_t S) {
  if (o < R) return false;
  bool eq = true;
  for (uint32_t k = 0; k < S; k += 4) eq = eq && (src.word(o + R - S + k) == src.word(o - S + k));
  return eq;
}

// chunk of thread t in a member of nb bytes: right-aligned frames of `chunk` bytes, so
// only leading threads are empty / partial (leading zero bytes leave a zero CRC
// register unchanged, which keeps the combination tree uniform)
__device__ __forceinline__ void gz_chunk(uint32_t nb, uint32_t chunk, int t, uint32_t& lo,
                                         uint32_t& hi) {
  const int64_t h = (int64_t)nb - (int64_t)(GZ_THREADS - 1 - t) * chunk;
  if (h <= 0) {
    lo = hi = 0;
    return;
  }
  hi = (uint32_t)h;
  lo = h > (int64_t)chunk ? (uint32_t)(h - chunk) : 0u;
}

__device__ __forceinline__ uint32_t gz_mat_apply(const uint32_t* m, uint32_t v) {
  uint32_t r = 0;
#pragma unroll
  for (int b = 0; b < 32; b++) r ^= ((v >> b) & 1u) ? m[b] : 0u;
  return r;
}

__global__ void __launch_bounds__(GZ_THREADS)
    gz_size_kernel(const unsigned char* __restrict__ data, uint64_t nbytes, uint32_t member_bytes,
                   const GzDev* __restrict__ T, uint32_t* __restrict__ thread_bits,
                   uint32_t* __restrict__ member_bits, uint32_t* __restrict__ member_crc) {
  // s_crc[k][b]: CRC register after byte b followed by k zero bytes (slicing by 4: the four
  // look-ups of a word are independent, where the byte-wise form is one dependent chain)
  __shared__ uint32_t s_len[256], s_crc[4][256], s_mat[6][32], s_red[GZ_THREADS / 32],
      s_wcrc[GZ_THREADS / 32];
  extern __shared__ __align__(16) uint32_t gz_stage_mem[];
  const int t = threadIdx.x;
  s_len[t] = T->entry[t] >> 16;
  s_crc[0][t] = T->crc_tab[t];
  if (t < 192) s_mat[t / 32][t % 32] = T->crc_mat[t / 32][t % 32];
  __syncthreads();
  {
    uint32_t c = s_crc[0][t];
#pragma unroll
    for (int k = 1; k < 4; k++) {
      c = (c >> 8) ^ s_crc[0][c & 255u];
      s_crc[k][t] = c;
    }
  }
  const uint64_t m = blockIdx.x;
  const uint64_t base = m * member_bytes;
  const uint32_t nb = (uint32_t)min((uint64_t)member_bytes, nbytes - base);
  uint32_t lo, hi;
  gz_chunk(nb, member_bytes / GZ_THREADS, t, lo, hi);
  uint32_t bits = 0, crc = 0;
  GzSrc src{reinterpret_cast<const uint32_t*>(data + base), nullptr, 0u, 0u};
  if (nb == member_bytes && member_bytes <= GZ_STAGE_MAX)
    src = gz_stage(data + base, member_bytes, gz_stage_mem);  // ends in __syncthreads()
  else
    __syncthreads();
  const uint32_t R = T->rec_bytes, S = T->tail_bytes, mbits = T->match_nbits;
  // chunks are whole records (or multiples of 4 bytes in plain mode)
  for (uint32_t r = lo; r < hi; r += R ? R : (hi - lo)) {
    const uint32_t rend = R ? r + R : hi;
    uint32_t lit_end = rend;
    if (S && gz_tail_match(src, r, R, S)) {
      lit_end -= S;
      bits += mbits;
    }
    for (uint32_t o = r; o < rend; o += 4) {
      const uint32_t w = src.word(o);
      if (o < lit_end)
        bits += s_len[w & 255u] + s_len[(w >> 8) & 255u] + s_len[(w >> 16) & 255u] + s_len[w >> 24];
      const uint32_t x = crc ^ w;
      crc = s_crc[3][x & 255u] ^ s_crc[2][(x >> 8) & 255u] ^ s_crc[1][(x >> 16) & 255u] ^
            s_crc[0][x >> 24];
    }
  }
  thread_bits[m * GZ_THREADS + t] = bits;
  // total bits of the member
  uint32_t s = bits;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  // CRC of the concatenation: c = M_j(left) ^ right, level by level
  const int lane = t & 31, warp = t >> 5;
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const uint32_t left = __shfl_up_sync(0xffffffffu, crc, 1 << j);
    if (((lane + 1) & ((2 << j) - 1)) == 0) crc = gz_mat_apply(s_mat[j], left) ^ crc;
  }
  if (lane == 0) s_red[warp] = s;
  if (lane == 31) s_wcrc[warp] = crc;
  __syncthreads();
  if (t == 0) {
    uint32_t tot = 0, c = 0;
    for (int w = 0; w < GZ_THREADS / 32; w++) {
      tot += s_red[w];
      c = gz_mat_apply(s_mat[5], c) ^ s_wcrc[w];
    }
    member_bits[m] = tot;
    const bool last = base + nb >= nbytes;
    member_crc[m] = c ^ (last ? T->init_last : T->init_full) ^ 0xffffffffu;
  }
}

// member m: 10-byte gzip header + FEXTRA (XLEN = 8: "KD", 4, member size) + k filler
// characters + NUL + deflate + CRC32 + ISIZE, k chosen so the member is a multiple of 4 bytes
constexpr uint32_t GZ_HEAD = 21u;  // bytes before the filler: 10 + 2 (XLEN) + 8 (subfield) + NUL
__device__ __forceinline__ uint32_t gz_member_size(uint32_t bits, uint32_t hdr_bits,
                                                   uint32_t eob_len, uint32_t& k) {
  const uint32_t D = (hdr_bits + bits + eob_len + 7) / 8;
  k = (4u - ((GZ_HEAD + 8u + D) & 3u)) & 3u;
  return GZ_HEAD + 8u + k + D;
}

__global__ void __launch_bounds__(GZ_THREADS)
    gz_scan_kernel(const uint32_t* __restrict__ member_bits, uint64_t n_members,
                   const GzDev* __restrict__ T, uint64_t* __restrict__ member_off,
                   uint64_t* __restrict__ total) {
  __shared__ uint64_t wsum[GZ_THREADS / 32];
  __shared__ uint64_t carry_s;
  const uint32_t hb = T->hdr_bits, eob = T->entry[256] >> 16;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint64_t b0 = 0; b0 < n_members; b0 += GZ_THREADS) {
    const uint64_t i = b0 + threadIdx.x;
    uint32_t k;
    const uint64_t v = i < n_members ? gz_member_size(member_bits[i], hb, eob, k) : 0;
    uint64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint64_t woff = 0, tot = 0;
    for (int w = 0; w < GZ_THREADS / 32; w++) {
      if (w < warp) woff += wsum[w];
      tot += wsum[w];
    }
    const uint64_t carry = carry_s;
    if (i < n_members) member_off[i] = carry + woff + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry_s;
}

__device__ __forceinline__ void gz_or_byte(uint32_t* words, uint32_t byte_pos, uint32_t v) {
  atomicOr(words + (byte_pos >> 2), v << (8 * (byte_pos & 3u)));
}

__global__ void __launch_bounds__(GZ_THREADS)
    gz_encode_kernel(const unsigned char* __restrict__ data, uint64_t nbytes, uint32_t member_bytes,
                     const GzDev* __restrict__ T, const uint32_t* __restrict__ thread_bits,
                     const uint32_t* __restrict__ member_bits,
                     const uint32_t* __restrict__ member_crc,
                     const uint64_t* __restrict__ member_off, uint64_t capacity,
                     unsigned char* __restrict__ out) {
  __shared__ uint32_t s_ent[257], s_scan[GZ_THREADS / 32];
  extern __shared__ __align__(16) uint32_t gz_stage_mem[];
  const int t = threadIdx.x;
  s_ent[t] = T->entry[t];
  if (t == 0) s_ent[256] = T->entry[256];
  const uint64_t m = blockIdx.x;
  const uint64_t base = m * member_bytes;
  const uint32_t nb = (uint32_t)min((uint64_t)member_bytes, nbytes - base);
  const uint32_t hb = T->hdr_bits, eob_len = T->entry[256] >> 16;
  uint32_t k;
  const uint32_t msize = gz_member_size(member_bits[m], hb, eob_len, k);
  const uint64_t off = member_off[m];
  if (off + msize > capacity) return;  // reported by the host from the scan total
  uint32_t* words = reinterpret_cast<uint32_t*>(out + off);  // off % 4 == 0
  // the member's bytes start from zero (headers, boundary words and the trailer are OR-ed in)
  for (uint32_t i = t; i < msize / 4; i += GZ_THREADS) words[i] = 0u;
  GzSrc src{reinterpret_cast<const uint32_t*>(data + base), nullptr, 0u, 0u};
  if (nb == member_bytes && member_bytes <= GZ_STAGE_MAX)
    src = gz_stage(data + base, member_bytes, gz_stage_mem);  // ends in __syncthreads()
  // exclusive prefix of the threads' bit counts
  const uint32_t mybits = thread_bits[m * GZ_THREADS + t];
  const int lane = t & 31, warp = t >> 5;
  uint32_t inc = mybits;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) s_scan[warp] = inc;
  __syncthreads();
  uint32_t woff = 0;
  for (int w = 0; w < warp; w++) woff += s_scan[w];
  const uint32_t B0 = 8u * (GZ_HEAD + k);  // first deflate bit
  const uint32_t start = B0 + hb + woff + inc - mybits;

  // gzip header (thread 0), block header bits (threads 0..), trailer (thread 1)
  if (t == 0) {
    atomicOr(words + 0, 0x0c088b1fu);  // ID1 ID2 CM=8 FLG=FEXTRA|FNAME; MTIME = 0
    gz_or_byte(words, 9, 0xffu);       // XFL = 0, OS = 255 (unknown)
    gz_or_byte(words, 10, 8u);         // XLEN = 8
    gz_or_byte(words, 12, 0x4bu);      // subfield "KD",
    gz_or_byte(words, 13, 0x44u);
    gz_or_byte(words, 14, 4u);         // 4 bytes: the size of this member
    atomicOr(words + 4, msize);
    for (uint32_t c = 0; c < k; c++) gz_or_byte(words, 20 + c, 0x6bu);  // filler 'k', then NUL
  }
  {
    const uint32_t s = B0 & 31u, nw = (s + hb + 31u) / 32u, hw = (hb + 31u) / 32u;
    for (uint32_t j = t; j < nw; j += GZ_THREADS) {
      uint32_t v = j < hw ? T->hdr[j] << s : 0u;
      if (s && j > 0) v |= T->hdr[j - 1] >> (32u - s);
      if (v) atomicOr(words + (B0 >> 5) + j, v);
    }
  }
  if (t == 1) {
    const uint32_t tb = (B0 + hb + member_bits[m] + eob_len + 7u) / 8u;
    const uint32_t crc = member_crc[m];
    for (int c = 0; c < 4; c++) {
      gz_or_byte(words, tb + c, (crc >> (8 * c)) & 255u);
      gz_or_byte(words, tb + 4 + c, (nb >> (8 * c)) & 255u);
    }
  }
  // codes of this thread's chunk; interior words are exclusively this thread's
  uint32_t lo, hi;
  gz_chunk(nb, member_bytes / GZ_THREADS, t, lo, hi);
  uint64_t acc = 0;
  uint32_t nbit = start & 31u;
  uint32_t* wp = words + (start >> 5);
  bool first = true;
  auto emit_bits = [&](uint32_t v, uint32_t n) {  // n <= 32 bits, LSB first
    acc |= (uint64_t)v << nbit;
    nbit += n;
    if (nbit >= 32u) {
      if (first) {
        atomicOr(wp, (uint32_t)acc);
        first = false;
      } else {
        *wp = (uint32_t)acc;
      }
      wp++;
      acc >>= 32;
      nbit -= 32u;
    }
  };
  auto emit = [&](uint32_t e) { emit_bits(e & 0xffffu, e >> 16); };
  const uint32_t R = T->rec_bytes, S = T->tail_bytes;
  const uint32_t mcode = T->match_bits, mbits = T->match_nbits;
  for (uint32_t r = lo; r < hi; r += R ? R : (hi - lo)) {
    uint32_t lit_end = R ? r + R : hi;
    const bool match = S && gz_tail_match(src, r, R, S);
    if (match) lit_end -= S;
    for (uint32_t o = r; o < lit_end; o += 4) {
      const uint32_t w = src.word(o);
      emit(s_ent[w & 255u]);
      emit(s_ent[(w >> 8) & 255u]);
      emit(s_ent[(w >> 16) & 255u]);
      emit(s_ent[w >> 24]);
    }
    if (match) emit_bits(mcode, mbits);
  }
  if (t == GZ_THREADS - 1) emit(s_ent[256]);  // end of block
  if (nbit) atomicOr(wp, (uint32_t)acc);
}

}  // namespace kdsb

using namespace kdsb;

extern "C" {

static int gz_check_records(uint32_t member_bytes, uint32_t rec_bytes, uint32_t tail_bytes,
                            uint64_t nbytes) {
  if (member_bytes == 0 || member_bytes % (GZ_THREADS * 4) != 0 || nbytes % 4 != 0 ||
      member_bytes > (1u << 24)) {
    set_error("gzip: member_bytes must be a multiple of %d (<= 2^24) and nbytes of 4",
              GZ_THREADS * 4);
    return 1;
  }
  if (rec_bytes == 0) return 0;
  if (rec_bytes % 4 || tail_bytes % 4 || tail_bytes > rec_bytes || rec_bytes > 32768 ||
      (tail_bytes && (tail_bytes < 4 || tail_bytes > 256)) ||
      member_bytes % (GZ_THREADS * rec_bytes) != 0 || nbytes % rec_bytes != 0) {
    set_error("gzip: record mode needs rec_bytes %% 4 == 0 (<= 32768), tail_bytes in 4..256 "
              "(%% 4 == 0), member_bytes a multiple of %d records and whole records",
              GZ_THREADS);
    return 1;
  }
  return 0;
}

int kdsb200_gz_histogram(const void* d_data, uint64_t nbytes, uint32_t member_bytes,
                         uint32_t rec_bytes, uint32_t tail_bytes, uint64_t* d_hist258,
                         void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  KDSB_CUDA(cudaMemsetAsync(d_hist258, 0, 258 * sizeof(uint64_t), st));
  if (nbytes < 4) return 0;
  if (gz_check_records(member_bytes, rec_bytes, tail_bytes, nbytes)) return 1;
  const uint64_t units = nbytes / (rec_bytes ? rec_bytes : 4);
  const unsigned blocks = (unsigned)std::min<uint64_t>((units + 255) / 256, (uint64_t)sm_count() * 8);
  gz_hist_kernel<<<blocks, 256, 0, st>>>((const unsigned char*)d_data, nbytes, member_bytes,
                                         rec_bytes, tail_bytes, (unsigned long long*)d_hist258);
  KDSB_CUDA(cudaGetLastError());
  return 0;
}

int kdsb200_gz_build_table(const uint64_t* h_hist258, uint32_t rec_bytes, uint32_t tail_bytes,
                           kdsb200_gz_table* out) {
  memset(out, 0, sizeof(*out));
  // RFC 1951 3.2.5: length / distance symbols with their base values and extra bits
  static const uint16_t lbase[29] = {3,  4,  5,  6,  7,  8,  9,  10, 11,  13,  15,  17,  19,  23, 27,
                                     31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2,
                                   2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  static const uint16_t dbase[30] = {1,   2,   3,   4,   5,   7,    9,    13,   17,   25,
                                     33,  49,  65,  97,  129, 193,  257,  385,  513,  769,
                                     1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2,  2,  3,  3,  4,  4,  5,  5,  6,
                                   6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  const bool use_match = rec_bytes && tail_bytes >= 3 && tail_bytes <= 258 &&
                         rec_bytes <= 32768 && h_hist258[257] > 0;
  int lsym = 0, dsym = 0;
  if (use_match) {
    while (lsym < 28 && lbase[lsym + 1] <= tail_bytes) lsym++;
    while (dsym < 29 && dbase[dsym + 1] <= rec_bytes) dsym++;
  }
  const int nlit = use_match ? 257 + lsym + 1 : 257;  // literal/length codes sent
  std::vector<uint64_t> freq(nlit, 0);
  for (int i = 0; i < 256; i++) freq[i] = h_hist258[i] + 1;  // every byte stays encodable
  freq[256] = 1;                                             // end of block
  if (use_match) freq[257 + lsym] = h_hist258[257];
  std::vector<uint8_t> len(nlit);
  std::vector<uint32_t> code(nlit);
  huffman_lengths(freq.data(), nlit, 15, len.data());
  canonical_codes(len.data(), nlit, code.data());
  memcpy(out->len, len.data(), 257);
  memcpy(out->code, code.data(), 257 * sizeof(uint32_t));
  // code lengths of the literal/length codes, then of the distance codes: either one
  // code of length 0 ("no distance codes", RFC 1951 3.2.7) or codes 0..dsym with only
  // the last one in use (a single distance code is sent with length 1)
  std::vector<uint8_t> seq(len);
  if (use_match) {
    for (int i = 0; i < dsym; i++) seq.push_back(0);
    seq.push_back(1);
  } else {
    seq.push_back(0);
  }
  const int ndist = use_match ? dsym + 1 : 1;
  uint64_t clfreq[19] = {0};
  for (uint8_t v : seq) clfreq[v]++;
  uint8_t cllen[19];
  uint32_t clcode[19];
  huffman_lengths(clfreq, 19, 7, cllen);
  canonical_codes(cllen, 19, clcode);
  static const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  int hclen = 19;
  while (hclen > 4 && cllen[order[hclen - 1]] == 0) hclen--;
  BitWriter bw(out->hdr, 64 * 32);
  bool ok = bw.put(1, 1) && bw.put(2, 2) && bw.put(nlit - 257, 5) && bw.put(ndist - 1, 5) &&
            bw.put(hclen - 4, 4);
  for (int i = 0; i < hclen && ok; i++) ok = bw.put(cllen[order[i]], 3);
  for (size_t i = 0; i < seq.size() && ok; i++) ok = bw.put(clcode[seq[i]], cllen[seq[i]]);
  if (!ok) {
    set_error("gz_build_table: block header does not fit %d bits", 64 * 32);
    return 1;
  }
  out->hdr_bits = bw.nbits;
  if (use_match) {
    // <length code><length extra bits><distance code: the single 1-bit code 0><extra bits>
    uint32_t words[2] = {0, 0};
    BitWriter mw(words, 64);
    mw.put(code[257 + lsym], len[257 + lsym]);
    mw.put(tail_bytes - lbase[lsym], lext[lsym]);
    mw.put(0, 1);
    mw.put(rec_bytes - dbase[dsym], dext[dsym]);
    if (mw.nbits <= 32) {
      out->rec_bytes = rec_bytes;
      out->tail_bytes = tail_bytes;
      out->match_bits = words[0];
      out->match_nbits = mw.nbits;
    } else {
      set_error("gz_build_table: match code longer than 32 bits");
      return 1;
    }
  }
  return 0;
}

uint32_t kdsb200_gz_member_bytes(void) { return 36u * 2048u; }

uint64_t kdsb200_gz_scratch_bytes(uint64_t nbytes, uint32_t member_bytes) {
  const uint64_t nm = (nbytes + member_bytes - 1) / member_bytes;
  // table image | thread_bits | member_bits | member_crc | member_off | total
  return 8192 + nm * (GZ_THREADS * 4 + 4 + 4 + 8) + 64;
}

int kdsb200_gz_compress(const void* d_data, uint64_t nbytes, uint32_t member_bytes,
                        const kdsb200_gz_table* h_table, void* d_scratch, void* d_out,
                        uint64_t out_capacity, uint64_t* d_total, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (nbytes == 0) {
    KDSB_CUDA(cudaMemsetAsync(d_total, 0, sizeof(uint64_t), st));
    return 0;
  }
  if (gz_check_records(member_bytes, h_table->rec_bytes, h_table->tail_bytes, nbytes)) return 1;
  if ((reinterpret_cast<uintptr_t>(d_data) & 3) || (reinterpret_cast<uintptr_t>(d_out) & 3) ||
      (reinterpret_cast<uintptr_t>(d_scratch) & 7) || (out_capacity & 3)) {
    set_error("gz_compress: buffers must be 4-byte aligned (scratch 8), capacity a multiple of 4");
    return 1;
  }
  const uint64_t nm = (nbytes + member_bytes - 1) / member_bytes;
  if (nm > 0x7fffffffull) {
    set_error("gz_compress: too many members");
    return 1;
  }
  // device image of the table (pinned staging is not needed for 8 KB)
  static thread_local GzDev img;
  for (int i = 0; i < 257; i++) img.entry[i] = ((uint32_t)h_table->len[i] << 16) | h_table->code[i];
  img.hdr_bits = h_table->hdr_bits;
  img.rec_bytes = h_table->rec_bytes;
  img.tail_bytes = h_table->tail_bytes;
  img.match_bits = h_table->match_bits;
  img.match_nbits = h_table->match_nbits;
  memcpy(img.hdr, h_table->hdr, sizeof(img.hdr));
  for (uint32_t n = 0; n < 256; n++) {
    uint32_t c = n;
    for (int k = 0; k < 8; k++) c = (c >> 1) ^ ((c & 1u) ? CRC_POLY : 0u);
    img.crc_tab[n] = c;
  }
  const uint32_t chunk = member_bytes / GZ_THREADS;
  for (int j = 0; j < 6; j++) crc_advance_matrix((uint64_t)chunk << j, img.crc_mat[j]);
  uint32_t mat[32];
  crc_advance_matrix(member_bytes, mat);
  img.init_full = mat_apply(mat, 0xffffffffu);
  crc_advance_matrix(nbytes - (nm - 1) * member_bytes, mat);
  img.init_last = mat_apply(mat, 0xffffffffu);
  static_assert(sizeof(GzDev) <= 8192, "table image must fit its scratch slot");

  unsigned char* sc = (unsigned char*)d_scratch;
  GzDev* dT = (GzDev*)sc;
  uint32_t* thread_bits = (uint32_t*)(sc + 8192);
  uint32_t* member_bits = thread_bits + nm * GZ_THREADS;
  uint32_t* member_crc = member_bits + nm;
  uint64_t* member_off = (uint64_t*)(sc + 8192 + ((nm * (GZ_THREADS * 4 + 8) + 7) / 8) * 8);
  KDSB_CUDA(cudaMemcpyAsync(dT, &img, sizeof(GzDev), cudaMemcpyHostToDevice, st));
  KDSB_CUDA(cudaStreamSynchronize(st));  // `img` is reused by the next call of this thread
  // full members are staged in shared memory (chunks at an odd word stride)
  size_t stage = 0;
  if (member_bytes <= GZ_STAGE_MAX) {
    const uint32_t cw = member_bytes / (4u * GZ_THREADS);
    stage = (size_t)GZ_THREADS * (cw + ((cw & 1u) ? 0u : 1u)) * 4;
  }
  static size_t attr_set = 0;
  if (stage > attr_set) {
    KDSB_CUDA(cudaFuncSetAttribute(gz_size_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)stage));
    KDSB_CUDA(cudaFuncSetAttribute(gz_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)stage));
    attr_set = stage;
  }
  gz_size_kernel<<<(unsigned)nm, GZ_THREADS, stage, st>>>((const unsigned char*)d_data, nbytes,
                                                          member_bytes, dT, thread_bits,
                                                          member_bits, member_crc);
  KDSB_CUDA(cudaGetLastError());
  gz_scan_kernel<<<1, GZ_THREADS, 0, st>>>(member_bits, nm, dT, member_off, d_total);
  KDSB_CUDA(cudaGetLastError());
  // (no memset of d_out: every member zeroes its own bytes before OR-ing into them)
  gz_encode_kernel<<<(unsigned)nm, GZ_THREADS, stage, st>>>(
      (const unsigned char*)d_data, nbytes, member_bytes, dT, thread_bits, member_bits, member_crc,
      member_off, out_capacity, (unsigned char*)d_out);
  KDSB_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
