// zlib streams of index planes on the device.
//
// The reference packs a frame as b'RHCCQ' + len + zlib9(pickle{.., 'i': zlib9(index bytes), ..})
// (/root/reference/encoder/compression/compression.py:151-220) and reads it back with zlib.decompress
// (/root/reference/decoder/uncompression/uncompression.py:58-92), which accepts any valid zlib stream.  zlib level 9
// of a full-HD index plane is 120 ms of a host core — seven times the whole encode for a batch —
// so this file produces the 'i' stream where the index plane already is.  The bytes differ from zlib's (only
// the host writer at level 9 is byte-identical to the reference); what is guaranteed, and tested, is that
// zlib.decompress returns the index bytes.
//
// Format produced (RFC 1950 / 1951): 78 01 | per 4 KiB chunk of the input: one block with the fixed Huffman
// code, then an empty stored block (00 00 00 FF FF after padding: zlib's "sync flush"), which leaves every chunk's
// output byte-aligned and independent of its neighbours | 03 00 (empty final block) | Adler-32, big-endian.
// Inside a chunk one thread walks the bytes: at each position the longest of five matches — against the previous
// index (distance = bytes per index), the index one row up (distance = bytes per row; index planes repeat
// vertically), its two neighbours (row -/+ one index) and the index two rows up — of at least 3 bytes becomes a
// length/distance pair, anything else a literal.  Matches
// may reach back across chunk borders (the window is the stream), never forward across them.
// A second kernel packs the chunks' outputs into one run per frame; the chunk sums for Adler-32 are combined on
// the host (two integers per chunk).
#include "rhccq_common.cuh"
#include "rhccq_kernels.h"

#define RHCCQ_DF_CHUNK 4096
#define RHCCQ_DF_SLOT (RHCCQ_DF_CHUNK + RHCCQ_DF_CHUNK / 8 + 16)   // worst case: 9 bits per literal + header, end, flush
#define RHCCQ_DF_THREADS 128

struct rhccq_bitw {
    uint8_t* out;
    unsigned long long acc;
    int nbits, pos;
};
__device__ __forceinline__ void rhccq_df_put(rhccq_bitw& w, unsigned v, int n) {           // LSB first
    w.acc |= (unsigned long long)v << w.nbits;
    w.nbits += n;
    while (w.nbits >= 8) { w.out[w.pos++] = (uint8_t)(w.acc & 0xffull); w.acc >>= 8; w.nbits -= 8; }
}
__device__ __forceinline__ unsigned rhccq_df_rev(unsigned v, int n) {                      // Huffman codes go MSB first
    unsigned r = 0;
    for (int i = 0; i < n; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
    return r;
}
__device__ __forceinline__ void rhccq_df_symbol(rhccq_bitw& w, int sym) {                  // fixed code of RFC 1951 3.2.6
    if (sym < 144) rhccq_df_put(w, rhccq_df_rev(0x30u + (unsigned)sym, 8), 8);
    else if (sym < 256) rhccq_df_put(w, rhccq_df_rev(0x190u + (unsigned)(sym - 144), 9), 9);
    else if (sym < 280) rhccq_df_put(w, rhccq_df_rev((unsigned)(sym - 256), 7), 7);
    else rhccq_df_put(w, rhccq_df_rev(0xC0u + (unsigned)(sym - 280), 8), 8);
}
__device__ __forceinline__ void rhccq_df_length(rhccq_bitw& w, int len) {                  // 3 .. 258
    if (len == 258) { rhccq_df_symbol(w, 285); return; }
    if (len <= 10) { rhccq_df_symbol(w, 254 + len); return; }
    // groups of four codes share an extra-bit count e = 1 .. 5: base length 3 + (4 << e) + code-in-group << e
    const int l = len - 3;
    int e = 1;
    while ((l >> (e + 2)) > 1) ++e;                                   // l in [4 << e, 8 << e)
    const int idx = (l - (4 << e)) >> e;                              // 0 .. 3
    rhccq_df_symbol(w, 261 + 4 * e + idx);
    rhccq_df_put(w, (unsigned)((l - (4 << e)) & ((1 << e) - 1)), e);
}
__device__ __forceinline__ void rhccq_df_distance(rhccq_bitw& w, int dist) {               // 1 .. 32768
    if (dist <= 4) { rhccq_df_put(w, rhccq_df_rev((unsigned)(dist - 1), 5), 5); return; }
    // pairs of codes share an extra-bit count e = 1 .. 13: base distance 1 + (2 << e) + code-in-pair << e
    const int d = dist - 1;
    int e = 1;
    while ((d >> (e + 1)) > 1) ++e;                                   // d in [2 << e, 4 << e)
    const int idx = (d - (2 << e)) >> e;                              // 0 .. 1
    rhccq_df_put(w, rhccq_df_rev((unsigned)(2 + 2 * e + idx), 5), 5);
    rhccq_df_put(w, (unsigned)((d - (2 << e)) & ((1 << e) - 1)), e);
}

// Length of the common prefix of a[0..) and a[-d..), at most `room` bytes.  Four bytes per step while eight bytes
// from both positions are inside the buffer (`left` bytes remain from a): two aligned words and a funnel shift give
// the word at any byte address.  `wide` is false when the buffer's base is not word-aligned.
__device__ __forceinline__ uint32_t rhccq_df_word(const uint8_t* a) {
#ifdef RHCCQ_HOST_EMU
    uint32_t v; memcpy(&v, a, 4); return v;
#else
    const uintptr_t ad = reinterpret_cast<uintptr_t>(a);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(ad & ~(uintptr_t)3);
    return __funnelshift_r(w[0], w[1], (unsigned)(ad & 3) * 8u);
#endif
}
__device__ __forceinline__ int rhccq_df_match(const uint8_t* a, int d, int room, long long left, bool wide) {
    int l = 0;
    if (wide) {
        while (l + 4 <= room && (long long)l + 8 <= left) {
            const uint32_t x = rhccq_df_word(a + l) ^ rhccq_df_word(a + l - d);
            if (x != 0u) return l + ((__ffs((int)x) - 1) >> 3);
            l += 4;
        }
    }
    while (l < room && a[l] == a[l - d]) ++l;
    return l;
}

// One thread per chunk.  frames: n_frames planes of frame_bytes bytes each, frame f at src + f * frame_stride.
// slots: n_frames * chunks_per_frame slots of RHCCQ_DF_SLOT bytes; slot_len: bytes written; sums: (sum of the
// chunk's bytes, sum of j * byte_j) for Adler-32.
__global__ void __launch_bounds__(RHCCQ_DF_THREADS)
rhccq_k_deflate_chunks(const uint8_t* __restrict__ src, long long frame_stride, int frame_bytes, int n_frames, int elem,
                       int row_bytes, uint8_t* __restrict__ slots, int* __restrict__ slot_len,
                       unsigned long long* __restrict__ sums) {
    const int cpf = (frame_bytes + RHCCQ_DF_CHUNK - 1) / RHCCQ_DF_CHUNK;
    const long long total = (long long)n_frames * cpf;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(c / cpf), ci = (int)(c % cpf);
        const uint8_t* p = src + (long long)f * frame_stride;
        const int s = ci * RHCCQ_DF_CHUNK, e = s + RHCCQ_DF_CHUNK < frame_bytes ? s + RHCCQ_DF_CHUNK : frame_bytes;
        rhccq_bitw w = {slots + c * RHCCQ_DF_SLOT, 0ull, 0, 0};
        rhccq_df_put(w, 2u, 3);                                       // BFINAL = 0, BTYPE = 01
        unsigned long long s1 = 0, s2 = 0;
        const bool use_row = row_bytes >= 1 && row_bytes <= 32768;
        const long long src_bytes = (long long)(n_frames - 1) * frame_stride + frame_bytes;
        const bool wide = (reinterpret_cast<uintptr_t>(src) & 3) == 0;
        int i = s;
        while (i < e) {
            const int room = e - i < 258 ? e - i : 258;
            int l1 = 0, l2 = 0;
            const long long left = src_bytes - ((long long)f * frame_stride + i);
            if (i >= elem) l1 = rhccq_df_match(p + i, elem, room, left, wide);
            if (use_row && i >= row_bytes && row_bytes != elem) l2 = rhccq_df_match(p + i, row_bytes, room, left, wide);
            int len = l1, dist = elem;
            if (l2 > l1) { len = l2; dist = row_bytes; }
            if (use_row && row_bytes > 2 * elem && row_bytes + elem <= 32768) {
                if (i >= row_bytes - elem) { const int l3 = rhccq_df_match(p + i, row_bytes - elem, room, left, wide); if (l3 > len) { len = l3; dist = row_bytes - elem; } }
                if (i >= row_bytes + elem) { const int l4 = rhccq_df_match(p + i, row_bytes + elem, room, left, wide); if (l4 > len) { len = l4; dist = row_bytes + elem; } }
            }
            if (use_row && 2 * row_bytes <= 32768 && i >= 2 * row_bytes) { const int l5 = rhccq_df_match(p + i, 2 * row_bytes, room, left, wide); if (l5 > len) { len = l5; dist = 2 * row_bytes; } }
            if (len >= 3) {
                rhccq_df_length(w, len);
                rhccq_df_distance(w, dist);
            } else {
                len = 1;
                rhccq_df_symbol(w, p[i]);
            }
            for (int j = 0; j < len; ++j) { const unsigned b = p[i + j]; s1 += b; s2 += (unsigned long long)(i + j - s) * b; }
            i += len;
        }
        rhccq_df_symbol(w, 256);                                      // end of block
        rhccq_df_put(w, 0u, 3);                                       // BFINAL = 0, BTYPE = 00: empty stored block ...
        if (w.nbits > 0) rhccq_df_put(w, 0u, 8 - w.nbits);            // ... after padding to a byte boundary
        rhccq_df_put(w, 0x0000u, 16);
        rhccq_df_put(w, 0xFFFFu, 16);
        slot_len[c] = w.pos;
        sums[2 * c] = s1;
        sums[2 * c + 1] = s2;
    }
}

// slot c's bytes -> out + offset[c] (offsets: the host's running sum of slot_len, frame by frame)
__global__ void __launch_bounds__(RHCCQ_DF_THREADS)
rhccq_k_deflate_pack(const uint8_t* __restrict__ slots, const int* __restrict__ slot_len, const long long* __restrict__ offset,
                     long long n_slots, uint8_t* __restrict__ out) {
    const int wpb = RHCCQ_NWARPS;
    for (long long c = (long long)blockIdx.x * wpb + RHCCQ_WARP; c < n_slots; c += (long long)gridDim.x * wpb) {
        const uint8_t* a = slots + c * RHCCQ_DF_SLOT;
        uint8_t* b = out + offset[c];
        const int n = slot_len[c];
        for (int j = RHCCQ_LANE; j < n; j += RHCCQ_WARP_SIZE) b[j] = a[j];
    }
}

extern "C" {

int rhccq_deflate_chunk_bytes(void) { return RHCCQ_DF_CHUNK; }
int rhccq_deflate_slot_bytes(void) { return RHCCQ_DF_SLOT; }

int rhccq_deflate_chunks(const uint8_t* src, long long frame_stride, int frame_bytes, int n_frames, int elem_bytes,
                         int row_bytes, uint8_t* slots, int32_t* slot_len, unsigned long long* sums, void* stream) {
    if (!src || !slots || !slot_len || !sums || frame_bytes < 0 || n_frames < 0 || elem_bytes < 1 || elem_bytes > 4 ||
        row_bytes < 0 || frame_stride < frame_bytes) {
        rhccq_set_error("rhccq_deflate_chunks: bad arguments");
        return -1;
    }
    const long long cpf = (frame_bytes + RHCCQ_DF_CHUNK - 1) / RHCCQ_DF_CHUNK, total = cpf * n_frames;
    if (total == 0) return 0;
    long long blocks = (total + RHCCQ_DF_THREADS - 1) / RHCCQ_DF_THREADS;
    const long long cap = (long long)rhccq_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    RHCCQ_LAUNCH(rhccq_k_deflate_chunks, (int)blocks, RHCCQ_DF_THREADS, 0, (cudaStream_t)stream, src, frame_stride, frame_bytes,
                 n_frames, elem_bytes, row_bytes, slots, slot_len, sums);
    return 0;
}

int rhccq_deflate_pack(const uint8_t* slots, const int32_t* slot_len, const long long* offsets, long long n_slots,
                       uint8_t* out, void* stream) {
    if (!slots || !slot_len || !offsets || !out || n_slots < 0) { rhccq_set_error("rhccq_deflate_pack: bad arguments"); return -1; }
    if (n_slots == 0) return 0;
    const int wpb = RHCCQ_DF_THREADS / 32;
    long long blocks = (n_slots + wpb - 1) / wpb;
    const long long cap = (long long)rhccq_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    RHCCQ_LAUNCH(rhccq_k_deflate_pack, (int)blocks, RHCCQ_DF_THREADS, 0, (cudaStream_t)stream, slots, slot_len, offsets, n_slots, out);
    return 0;
}

}  // extern "C"
