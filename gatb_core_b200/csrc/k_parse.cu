// k_parse.cu -- FASTA / FASTQ text -> packed 2-bit reads on the device.
//
// Replaces the per-record line scan of the reference's parser
//   BankFasta::Iterator::get_next_seq / buffered line reads     bank/impl/BankFasta.cpp:391-620
//   Data::ConvertASCII (nucleotide codes, validity)              tools/misc/api/Data.hpp:185-189
// (paths relative to /root/reference/gatb-core/src/gatb/) for plain-text files: the host only cuts the file into batches at
// record boundaries; lines, record boundaries, sequence offsets and the 2-bit packing are found on the device.
//   k_text_count_newlines / k_text_line_starts   two-pass compaction of the newline positions -> line starts
//   k_text_lines          one thread per line: kind (header / sequence / other), sequence characters of the line
//   (two scans: destination nucleotide of every line, record index of every header)
//   k_text_offsets        read_offsets_nt of the records (one per header line) appended to the context's offsets
//   k_text_pack           one thread per 32 output nucleotides: finds its line by binary search over the line destinations and
//                         walks the characters across line breaks; same codes and validity rule as k_pack_ascii
//   k_text_stats          sequences, nucleotides, shortest / longest record, sum of squared lengths (BankStats)
// FASTA: a record starts at a line whose first character is '>'; every other non-empty line up to the next header is sequence
// (multi-line records).  FASTQ: records of four lines (header, sequence, '+', qualities).  '\r' before '\n' is dropped.
#include "common.cuh"
#include "kernels.h"

#define TEXT_BLOCK 4096          // bytes per CTA of the newline passes

__global__ void __launch_bounds__(256) k_text_count_newlines (const char* __restrict__ text, uint64_t n, uint32_t* __restrict__ block_counts)
{
    const uint64_t b0 = (uint64_t)blockIdx.x * TEXT_BLOCK;
    uint32_t c = 0;
    for (uint32_t i = threadIdx.x * 16; i < TEXT_BLOCK; i += 256 * 16)
    {
        const uint64_t p = b0 + i;
        if (p + 16 <= n) { const uint4 v = *(const uint4*)(text + p); const uint32_t w[4] = { v.x, v.y, v.z, v.w };
                           #pragma unroll
                           for (int q = 0; q < 4; q++) c += __popc (__vcmpeq4 (w[q], 0x0A0A0A0Au)) >> 3; }
        else for (uint64_t q = p; q < n && q < p + 16; q++) c += (text[q] == '\n');
    }
    __shared__ uint32_t s[8];
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync (FULL_MASK, c, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads ();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int i = 0; i < 8; i++) t += s[i]; block_counts[blockIdx.x] = t; }
}
// line_start[1 + rank of the newline] = position after it (line_start[0] = 0 is set by the host)
__global__ void __launch_bounds__(256) k_text_line_starts (const char* __restrict__ text, uint64_t n, const uint64_t* __restrict__ block_off, uint64_t* __restrict__ line_start)
{
    const uint64_t b0 = (uint64_t)blockIdx.x * TEXT_BLOCK;
    __shared__ uint32_t s_warp[8];
    uint64_t base = block_off[blockIdx.x];
    for (uint32_t i0 = 0; i0 < TEXT_BLOCK; i0 += 256)
    {
        const uint64_t p = b0 + i0 + threadIdx.x;
        const bool nl = p < n && text[p] == '\n';
        const unsigned m = __ballot_sync (FULL_MASK, nl);
        if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = __popc (m);
        __syncthreads ();
        uint32_t before = 0, total = 0;
        for (int w = 0; w < 8; w++) { const uint32_t v = s_warp[w]; if (w < (int)(threadIdx.x >> 5)) before += v; total += v; }
        if (nl) line_start[1 + base + before + __popc (m & ((1u << (threadIdx.x & 31)) - 1))] = p + 1;
        base += total;
        __syncthreads ();
    }
}
// per line: sequence characters (0 for headers and ignored lines) and header flag
__global__ void k_text_lines (const char* __restrict__ text, uint64_t n, const uint64_t* __restrict__ line_start, uint64_t n_lines, int format,
                              uint32_t* __restrict__ seq_len, uint32_t* __restrict__ is_header)
{
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_lines) return;
    const uint64_t a = line_start[j];
    uint64_t e = (j + 1 < n_lines) ? line_start[j + 1] - 1 : n;                 // the last line may lack its newline
    if (j + 1 == n_lines && e > a && text[e - 1] == '\n') e--;
    if (e > a && text[e - 1] == '\r') e--;
    bool hdr, seq;
    if (format == 1) { hdr = (j & 3) == 0 && e > a; seq = (j & 3) == 1; }
    else             { const char c = e > a ? text[a] : 0; hdr = c == '>'; seq = e > a && c != '>' && c != ';'; }
    seq_len[j] = seq ? (uint32_t)(e - a) : 0u;
    is_header[j] = hdr ? 1u : 0u;
}
// offsets of the records of this batch: out[rec] = base + destination of the first sequence character after header line j
__global__ void k_text_offsets (const uint32_t* __restrict__ is_header, const uint64_t* __restrict__ line_dst, const uint64_t* __restrict__ line_rec,
                                uint64_t n_lines, uint64_t base, uint64_t* __restrict__ out, uint64_t n_recs, uint64_t total_nt)
{
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_lines && is_header[j]) out[line_rec[j]] = base + line_dst[j];
    if (j == 0) out[n_recs] = base + total_nt;
}
__global__ void __launch_bounds__(256) k_text_pack (const char* __restrict__ text, const uint64_t* __restrict__ line_start, const uint64_t* __restrict__ line_dst,
                                                    uint64_t n_lines, uint64_t total_nt, uint64_t base, uint32_t* words, uint32_t* nmask,
                                                    unsigned long long* n_invalid)
{
    unsigned long long bad = 0;
    if (total_nt)
    {
        const uint64_t g_first = base / 32, g_last = (base + total_nt - 1) / 32;
        for (uint64_t g = g_first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g <= g_last; g += (uint64_t)gridDim.x * blockDim.x)
        {
            uint64_t pos = g * 32 < base ? base : g * 32;                       // first stream position this thread writes
            const uint64_t pos_end = (g * 32 + 32 < base + total_nt) ? g * 32 + 32 : base + total_nt;
            // the line that holds batch nucleotide (pos - base): last line with line_dst <= it (empty lines share the next one's value)
            const uint64_t want = pos - base;
            uint64_t lo = 0, hi = n_lines;
            while (hi - lo > 1) { const uint64_t mid = (lo + hi) >> 1; if (line_dst[mid] <= want) lo = mid; else hi = mid; }
            uint64_t j = lo;
            uint64_t src = line_start[j] + (want - line_dst[j]);
            uint64_t left = ((j + 1 < n_lines) ? line_dst[j + 1] : total_nt) - want;   // characters left in this line
            uint32_t w0 = 0, w1 = 0, mk = 0;
            for (; pos < pos_end; pos++)
            {
                while (left == 0) { j++; src = line_start[j]; left = ((j + 1 < n_lines) ? line_dst[j + 1] : total_nt) - line_dst[j]; }
                const unsigned char c = (unsigned char)text[src];
                src++; left--;
                const int t = (int)(pos & 31);
                const uint32_t code = (c >> 1) & 3u;
                const bool ok = (c=='A'||c=='C'||c=='G'||c=='T'||c=='a'||c=='c'||c=='g'||c=='t');
                if (!ok) { mk |= 1u << t; bad++; }
                if (t < 16) w0 |= code << (2*t); else w1 |= code << (2*(t-16));
            }
            if (g == g_first || g == g_last) { atomicOr (&words[2*g], w0); atomicOr (&words[2*g+1], w1); atomicOr (&nmask[g], mk); }
            else { words[2*g] = w0; words[2*g+1] = w1; nmask[g] = mk; }
        }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync (FULL_MASK, bad, o);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd (n_invalid, bad);
}
// stats over records [first, first+n): [0] shortest [1] longest [2] sum of squared lengths (as double bits)
__global__ void k_text_stats (const uint64_t* __restrict__ offsets, uint64_t n, unsigned long long* out)
{
    unsigned long long mn = ~0ULL, mx = 0; double sq = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    {
        const unsigned long long len = offsets[i + 1] - offsets[i];
        mn = len < mn ? len : mn; mx = len > mx ? len : mx; sq += (double)len * (double)len;
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        const unsigned long long a = __shfl_xor_sync (FULL_MASK, mn, o), b = __shfl_xor_sync (FULL_MASK, mx, o);
        mn = a < mn ? a : mn; mx = b > mx ? b : mx; sq += __shfl_xor_sync (FULL_MASK, sq, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicMin (&out[0], mn); atomicMax (&out[1], mx); atomicAdd ((double*)&out[2], sq); }
}

cudaError_t launch_text_count_newlines (const LaunchCtx& L, const char* text, uint64_t n, uint32_t* block_counts)
{
    const unsigned blocks = (unsigned)((n + TEXT_BLOCK - 1) / TEXT_BLOCK);
    if (!blocks) return cudaSuccess;
    k_text_count_newlines<<<blocks, 256, 0, L.stream>>> (text, n, block_counts); (*L.launches)++;
    return cudaGetLastError ();
}
cudaError_t launch_text_line_starts (const LaunchCtx& L, const char* text, uint64_t n, const uint64_t* block_off, uint64_t* line_start)
{
    const unsigned blocks = (unsigned)((n + TEXT_BLOCK - 1) / TEXT_BLOCK);
    if (!blocks) return cudaSuccess;
    k_text_line_starts<<<blocks, 256, 0, L.stream>>> (text, n, block_off, line_start); (*L.launches)++;
    return cudaGetLastError ();
}
uint64_t text_blocks (uint64_t n) { return (n + TEXT_BLOCK - 1) / TEXT_BLOCK; }
cudaError_t launch_text_lines (const LaunchCtx& L, const char* text, uint64_t n, const uint64_t* line_start, uint64_t n_lines, int format,
                               uint32_t* seq_len, uint32_t* is_header)
{
    if (!n_lines) return cudaSuccess;
    k_text_lines<<<(unsigned)((n_lines + 255) / 256), 256, 0, L.stream>>> (text, n, line_start, n_lines, format, seq_len, is_header); (*L.launches)++;
    return cudaGetLastError ();
}
cudaError_t launch_text_offsets (const LaunchCtx& L, const uint32_t* is_header, const uint64_t* line_dst, const uint64_t* line_rec, uint64_t n_lines,
                                 uint64_t base, uint64_t* out, uint64_t n_recs, uint64_t total_nt)
{
    k_text_offsets<<<(unsigned)((n_lines + 255) / 256 + 1), 256, 0, L.stream>>> (is_header, line_dst, line_rec, n_lines, base, out, n_recs, total_nt); (*L.launches)++;
    return cudaGetLastError ();
}
cudaError_t launch_text_pack (const LaunchCtx& L, const char* text, const uint64_t* line_start, const uint64_t* line_dst, uint64_t n_lines,
                              uint64_t total_nt, uint64_t base, uint32_t* words, uint32_t* nmask, unsigned long long* n_invalid)
{
    if (!total_nt) return cudaSuccess;
    const uint64_t groups = (base + total_nt - 1) / 32 - base / 32 + 1;
    const uint64_t blocks = (groups + 255) / 256; const unsigned grid = (unsigned)(blocks < (uint64_t)L.sm_count * 32 ? blocks : (uint64_t)L.sm_count * 32);
    k_text_pack<<<grid, 256, 0, L.stream>>> (text, line_start, line_dst, n_lines, total_nt, base, words, nmask, n_invalid); (*L.launches)++;
    return cudaGetLastError ();
}
cudaError_t launch_text_stats (const LaunchCtx& L, const uint64_t* offsets, uint64_t n, unsigned long long* out)
{
    if (!n) return cudaSuccess;
    k_text_stats<<<L.sm_count * 4, 256, 0, L.stream>>> (offsets, n, out); (*L.launches)++;
    return cudaGetLastError ();
}
