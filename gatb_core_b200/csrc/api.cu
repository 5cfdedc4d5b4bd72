// api.cu -- the extern "C" boundary (include/gatb_gpu.h): context, buffers, stage orchestration.
// No CPU fallback anywhere: every compute entry point needs the CUDA device of its context.
#include "../../include/gatb_gpu.h"
#include "common.cuh"
#include "kernels.h"
#include "gatb_tables.h"
#include "repart_host.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <queue>
#include <thread>
#include <algorithm>

static std::string g_create_error;

// device buffer slots cached in the context so that repeated calls (bench steps) do not re-allocate
enum Slot { S_READS, S_OFFSETS, S_NMASK, S_COARSE, S_FINE, S_CURSORS, S_BINDESC, S_STATS, S_HISTO, S_COUNTERS,
            S_OVFLIST, S_GTABLE, S_REPART, S_BUCKETOF, S_BUCKETCNT, S_BUCKETOFF, S_SCAN, S_BIGLIST, S_SORTED, S_MISC, S_MISC2, S_TOTCUR, S_COARSEOFF, S_UNSORTED, S_RESMISC, S_DIR, S_OVFLIST2, S_BINOFF, S_PRESPLIT, S_SUBOFF, S_SUBCNT, S_ROUTECNT, S_NSLOTS };

struct gatb_gpu_ctx
{
    int device; int sm_count; cudaStream_t stream; uint64_t launches;
    std::string error;
    void* slot[S_NSLOTS]; size_t slot_cap[S_NSLOTS];
    cudaEvent_t ev[8]; cudaEvent_t kev[16];
    cudaStream_t copy_stream; cudaEvent_t cev[40];
    void* pinned; size_t pinned_cap;
    const uint16_t* repart_host_cached; uint64_t repart_bytes_cached;
    // streaming input (gatb_gpu_reads_*): reads pushed so far live in S_READS / S_OFFSETS / S_NMASK
    uint64_t push_nt, push_seqs, push_invalid;
    void* multi_buf; size_t multi_cap;          // merged host result of gatb_gpu_count_multi (first context)
};

static int fail (gatb_gpu_ctx* c, const char* fmt, ...)
{
    char buf[1024]; va_list ap; va_start (ap, fmt); vsnprintf (buf, sizeof(buf), fmt, ap); va_end (ap);
    if (c) c->error = buf; else g_create_error = buf;
    return 1;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail (ctx, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString (e_)); } while (0)

static int ensure (gatb_gpu_ctx* ctx, int s, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    if (ctx->slot_cap[s] >= bytes) return 0;
    if (ctx->slot[s]) { cudaStreamSynchronize (ctx->stream); cudaFree (ctx->slot[s]); ctx->slot[s] = 0; ctx->slot_cap[s] = 0; }
    size_t want = bytes + bytes / 16 + 256;                         // a little head-room against re-allocation
    cudaError_t e = cudaMalloc (&ctx->slot[s], want);
    if (e != cudaSuccess) { e = cudaMalloc (&ctx->slot[s], bytes); want = bytes; }
    if (e != cudaSuccess) return fail (ctx, "cudaMalloc of %zu bytes (slot %d) failed: %s", bytes, s, cudaGetErrorString (e));
    ctx->slot_cap[s] = want;
    return 0;
}
static void release (gatb_gpu_ctx* ctx, int s)
{ if (ctx->slot[s]) { cudaStreamSynchronize (ctx->stream); cudaFree (ctx->slot[s]); ctx->slot[s] = 0; ctx->slot_cap[s] = 0; } }

static void* pinned_ensure (gatb_gpu_ctx* ctx, size_t bytes)
{
    if (ctx->pinned_cap >= bytes) return ctx->pinned;
    if (ctx->pinned) { cudaStreamSynchronize (ctx->stream); cudaFreeHost (ctx->pinned); ctx->pinned = 0; ctx->pinned_cap = 0; }
    size_t want = bytes + bytes / 8 + 4096;
    if (cudaHostAlloc (&ctx->pinned, want, cudaHostAllocDefault) != cudaSuccess) { ctx->pinned = 0; return 0; }
    ctx->pinned_cap = want;
    return ctx->pinned;
}

static LaunchCtx lctx (gatb_gpu_ctx* ctx) { LaunchCtx L; L.stream = ctx->stream; L.sm_count = ctx->sm_count; L.launches = &ctx->launches; return L; }

extern "C" {

gatb_gpu_ctx* gatb_gpu_create (int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount (&n);
    if (e != cudaSuccess || n == 0) { fail (0, "no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString (e)); return 0; }
    if (device < 0 || device >= n) { fail (0, "device %d out of range (%d devices)", device, n); return 0; }
    if ((e = cudaSetDevice (device)) != cudaSuccess) { fail (0, "cudaSetDevice(%d): %s", device, cudaGetErrorString (e)); return 0; }
    gatb_gpu_ctx* ctx = new gatb_gpu_ctx ();
    ctx->device = device; ctx->launches = 0; ctx->pinned = 0; ctx->pinned_cap = 0; ctx->repart_host_cached = 0; ctx->repart_bytes_cached = 0;
    ctx->push_nt = ctx->push_seqs = ctx->push_invalid = 0; ctx->multi_buf = 0; ctx->multi_cap = 0;
    memset (ctx->slot, 0, sizeof(ctx->slot)); memset (ctx->slot_cap, 0, sizeof(ctx->slot_cap));
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties (&prop, device)) != cudaSuccess) { fail (0, "cudaGetDeviceProperties: %s", cudaGetErrorString (e)); delete ctx; return 0; }
    ctx->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags (&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) { fail (0, "cudaStreamCreate: %s", cudaGetErrorString (e)); delete ctx; return 0; }
    for (int i = 0; i < 8; i++) cudaEventCreate (&ctx->ev[i]);
    for (int i = 0; i < 16; i++) cudaEventCreate (&ctx->kev[i]);
    cudaStreamCreateWithFlags (&ctx->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 40; i++) cudaEventCreateWithFlags (&ctx->cev[i], cudaEventDisableTiming);
    return ctx;
}
void gatb_gpu_destroy (gatb_gpu_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice (ctx->device);
    cudaStreamSynchronize (ctx->stream);
    for (int s = 0; s < S_NSLOTS; s++) if (ctx->slot[s]) cudaFree (ctx->slot[s]);
    if (ctx->pinned) cudaFreeHost (ctx->pinned);
    free (ctx->multi_buf);
    for (int i = 0; i < 8; i++) cudaEventDestroy (ctx->ev[i]);
    for (int i = 0; i < 16; i++) cudaEventDestroy (ctx->kev[i]);
    for (int i = 0; i < 40; i++) cudaEventDestroy (ctx->cev[i]);
    cudaStreamDestroy (ctx->copy_stream);
    cudaStreamDestroy (ctx->stream);
    delete ctx;
}
const char* gatb_gpu_last_error (gatb_gpu_ctx* ctx) { return ctx ? ctx->error.c_str () : g_create_error.c_str (); }
void*    gatb_gpu_stream (gatb_gpu_ctx* ctx) { return (void*)ctx->stream; }
uint64_t gatb_gpu_kernel_launches (gatb_gpu_ctx* ctx) { return ctx->launches; }
int      gatb_gpu_sm_count (gatb_gpu_ctx* ctx) { return ctx->sm_count; }

void* gatb_gpu_malloc (gatb_gpu_ctx* ctx, uint64_t bytes)
{ cudaSetDevice (ctx->device); void* p = 0; if (cudaMalloc (&p, bytes ? bytes : 16) != cudaSuccess) { fail (ctx, "cudaMalloc(%llu) failed", (unsigned long long)bytes); return 0; } return p; }
void gatb_gpu_free (gatb_gpu_ctx* ctx, void* p) { cudaSetDevice (ctx->device); if (p) cudaFree (p); }
int gatb_gpu_memcpy_h2d (gatb_gpu_ctx* ctx, void* dst, const void* src, uint64_t bytes)
{ cudaSetDevice (ctx->device); CK (cudaMemcpyAsync (dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream)); CK (cudaStreamSynchronize (ctx->stream)); return 0; }
int gatb_gpu_memcpy_d2h (gatb_gpu_ctx* ctx, void* dst, const void* src, uint64_t bytes)
{ cudaSetDevice (ctx->device); CK (cudaMemcpyAsync (dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream)); CK (cudaStreamSynchronize (ctx->stream)); return 0; }
int gatb_gpu_synchronize (gatb_gpu_ctx* ctx) { cudaSetDevice (ctx->device); CK (cudaStreamSynchronize (ctx->stream)); return 0; }

int gatb_gpu_synth_reads_dev (gatb_gpu_ctx* ctx, uint64_t seed, uint64_t genome_len, uint64_t first_read, uint64_t n_reads, int L, uint8_t* d_packed)
{
    cudaSetDevice (ctx->device);
    if (genome_len < (uint64_t)L) return fail (ctx, "genome_len < read length");
    CK (launch_synth_reads (lctx (ctx), seed, genome_len, first_read, n_reads, L, d_packed));
    return 0;
}

int gatb_gpu_synth_zipf_dev (gatb_gpu_ctx* ctx, uint64_t seed, uint64_t n_species, const uint64_t* cdf, const uint64_t* genome_off,
                             uint64_t first_read, uint64_t n_reads, int L, uint8_t* d_packed)
{
    cudaSetDevice (ctx->device);
    if (n_species == 0 || !cdf || !genome_off) return fail (ctx, "synth_zipf: tables missing");
    for (uint64_t s = 0; s < n_species; s++) if (genome_off[s+1] - genome_off[s] < (uint64_t)L) return fail (ctx, "synth_zipf: a genome is shorter than the reads");
    if (ensure (ctx, S_MISC2, (2 * n_species + 1) * 8 + 64)) return 1;
    uint64_t* d_cdf = (uint64_t*)ctx->slot[S_MISC2]; uint64_t* d_off = d_cdf + n_species;
    CK (cudaMemcpyAsync (d_cdf, cdf, n_species * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK (cudaMemcpyAsync (d_off, genome_off, (n_species + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK (launch_synth_reads_zipf (lctx (ctx), seed, n_species, d_cdf, d_off, first_read, n_reads, L, d_packed));
    CK (cudaStreamSynchronize (ctx->stream));
    return 0;
}

int gatb_gpu_pack_ascii (gatb_gpu_ctx* ctx, const char* ascii, uint64_t n, uint8_t* packed_out, uint32_t* n_mask_out, uint64_t* n_invalid)
{
    cudaSetDevice (ctx->device);
    uint64_t groups = (n + 31) / 32;
    if (ensure (ctx, S_MISC, n + 64)) return 1;
    if (ensure (ctx, S_MISC2, groups * 12 + 64)) return 1;
    if (ensure (ctx, S_STATS, 64 * 8)) return 1;
    char* d_ascii = (char*)ctx->slot[S_MISC];
    uint32_t* d_words = (uint32_t*)ctx->slot[S_MISC2]; uint32_t* d_mask = d_words + 2 * groups;
    unsigned long long* d_bad = (unsigned long long*)ctx->slot[S_STATS];
    CK (cudaMemcpyAsync (d_ascii, ascii, n, cudaMemcpyHostToDevice, ctx->stream));
    CK (cudaMemsetAsync (d_bad, 0, 8, ctx->stream));
    CK (launch_pack_ascii (lctx (ctx), d_ascii, n, d_words, d_mask, d_bad));
    CK (cudaMemcpyAsync (packed_out, d_words, (n + 3) / 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (n_mask_out) CK (cudaMemcpyAsync (n_mask_out, d_mask, groups * 4, cudaMemcpyDeviceToHost, ctx->stream));
    unsigned long long bad = 0;
    CK (cudaMemcpyAsync (&bad, d_bad, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK (cudaStreamSynchronize (ctx->stream));
    if (n_invalid) *n_invalid = bad;
    return 0;
}

} // extern "C"

// =====================================================================================================================
//  DSK
// =====================================================================================================================
struct DevResult      // device-side arrays of one gatb_gpu_count_dev call (freed by gatb_gpu_result_free)
{
    void* lo; void* hi; void* cnt; void* histo; void* offs;
};

static int check_params (gatb_gpu_ctx* ctx, const gatb_gpu_params* p, const uint16_t* repart)
{
    if (!p) return fail (ctx, "params is NULL");
    if (p->kmer_size < 2 || p->kmer_size > 63) return fail (ctx, "kmer_size %d not supported (2..63: Kmer<32> and Kmer<64>)", p->kmer_size);
    if (p->minimizer_size < 2 || p->minimizer_size > 12 || p->minimizer_size >= p->kmer_size)      // (m = 1 has no "AA" mask: (m-2)*2 bits)
        return fail (ctx, "Bad values for kmer %d and minimizer %d", p->kmer_size, p->minimizer_size);     // Model.hpp:1014
    if (p->nb_partitions < 1 || p->nb_passes < 1) return fail (ctx, "nb_partitions and nb_passes must be >= 1");
    if ((uint64_t)p->nb_partitions * p->nb_passes > 65535) return fail (ctx, "too many partition keys");
    if ((uint64_t)p->nb_partitions * p->nb_passes > 1 && !repart) return fail (ctx, "a Repartitor table is required when nb_partitions*nb_passes > 1");
    if (p->minimizer_type != 0) return fail (ctx, "minimizer_type %d (frequency order) is not supported on the device path yet", p->minimizer_type);
    if (p->histo_max < 1) return fail (ctx, "histo_max must be >= 1");
    return 0;
}

static int pick_device_m (int k, uint64_t nbins_fine)
{
    // smallest m whose canonical m-mer space is >= 16x the number of device bins, within [8, 15] and < k
    int m = 8;
    while (m < 15 && (1ULL << (2*m - 1)) < 16 * nbins_fine) m++;
    if (m > k - 1) m = k - 1;
    if (m < 1) m = 1;
    return m;
}

// total number of k-mer positions of a batch of reads
__global__ void k_count_kmers (const uint64_t* offsets, uint64_t n_reads, int k, unsigned long long* out /* [0] kmers [1] nt [2] maxlen */)
{
    unsigned long long km = 0, nt = 0, mx = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_reads; i += (uint64_t)gridDim.x * blockDim.x)
    {
        unsigned long long len = offsets[i+1] - offsets[i];
        nt += len; if (len >= (unsigned long long)k) km += len - k + 1; if (len > mx) mx = len;
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) { km += __shfl_xor_sync (FULL_MASK, km, o); nt += __shfl_xor_sync (FULL_MASK, nt, o); unsigned long long y = __shfl_xor_sync (FULL_MASK, mx, o); mx = y > mx ? y : mx; }
    if ((threadIdx.x & 31) == 0) { atomicAdd (&out[0], km); atomicAdd (&out[1], nt); atomicMax (&out[2], mx); }
}

// tot[b] = sum over sources of min(cursor_s[b], cap)
struct CursorList { const uint32_t* cur[GATB_GPU_MAX_SOURCES]; int n; };
// fine bins of the listed coarse bins: out[i << fine_bits | f] = list[i] << fine_bits | f
__global__ void k_expand_bins (const uint32_t* list, uint32_t n, int fine_bits, uint32_t* out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ((uint64_t)n << fine_bits)) out[i] = (list[i >> fine_bits] << fine_bits) | (uint32_t)(i & ((1u << fine_bits) - 1));
}
__global__ void k_sum_cursors (CursorList C, uint32_t nb1, uint32_t cap, uint32_t* tot, uint32_t* max_tot)
{
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t t = 0;
    if (b < nb1)
    {
        for (int s = 0; s < C.n; s++) t += min (C.cur[s][b], cap);
        tot[b] = t;
    }
    t = __reduce_max_sync (FULL_MASK, t);
    if ((threadIdx.x & 31) == 0 && t) atomicMax (max_tot, t);
}

static int workload_size (gatb_gpu_ctx* ctx, const gatb_gpu_params* p, const uint64_t* d_offsets, uint64_t n_reads,
                          uint64_t* total_kmers, uint64_t* total_nt, uint64_t* max_len)
{
    const int k = p->kmer_size;
    if (!d_offsets)
    {
        if (p->read_len <= 0) return fail (ctx, "read_offsets_nt is NULL and read_len <= 0");
        *total_nt = n_reads * (uint64_t)p->read_len; *max_len = p->read_len;
        *total_kmers = p->read_len >= k ? n_reads * (uint64_t)(p->read_len - k + 1) : 0;
        return 0;
    }
    if (ensure (ctx, S_STATS, 64 * 8)) return 1;
    unsigned long long* d_stats = (unsigned long long*)ctx->slot[S_STATS];
    CK (cudaMemsetAsync (d_stats + 32, 0, 3 * 8, ctx->stream));
    if (n_reads) { k_count_kmers<<<ctx->sm_count * 4, 256, 0, ctx->stream>>> (d_offsets, n_reads, k, d_stats + 32); ctx->launches++; }
    unsigned long long h[3];
    CK (cudaMemcpyAsync (h, d_stats + 32, 3 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK (cudaStreamSynchronize (ctx->stream));
    *total_kmers = h[0]; *total_nt = h[1]; *max_len = h[2];
    return 0;
}

// k <= 31 can count straight out of the coarse bins (k2_fused.cu, opt-in) instead of the fine-split pipeline
static bool path_fused (const gatb_gpu_params* p)
{ return p->kmer_size < 32 && (p->path_flags & GATB_PATH_FUSED) && !(p->path_flags & GATB_PATH_K2B_MASK); }

// ---- geometry of the device binning: must be identical on every rank of a multi-GPU run ----------------------------
static int plan_geometry (gatb_gpu_ctx* ctx, const gatb_gpu_params* p, uint64_t total_kmers, uint64_t n_reads, int n_ranks, gatb_gpu_geometry* g)
{
    const int k = p->kmer_size, W = (k < 32) ? 1 : 2;
    const bool fused = path_fused (p);
    int table_log2 = p->table_log2 > 0 ? p->table_log2 : (fused ? 13 : k2b_default_table_log2 (W, p->path_flags));
    if (table_log2 < 5 || table_log2 > 13) return fail (ctx, "table_log2 must be in [5,13]");
    if (n_ranks < 1 || n_ranks > GATB_GPU_MAX_RANKS) return fail (ctx, "n_ranks must be in [1,%d]", GATB_GPU_MAX_RANKS);
    const uint64_t T = 1ULL << table_log2;
    if (p->bin_load_pct < 0 || p->bin_load_pct > 400) return fail (ctx, "bin_load_pct must be in [0,400]");
    // k <= 31: a coarse bin is planned for 64 counting bins (measured best on B200: the dedup split stages a ~3300-record coarse bin
    // in shared memory, several CTAs per SM) and carries 512 fine ids: the dedup split merges consecutive ids into counting bins of
    // even load (k2a_dedup_split, "adaptive bins").  One more id bit per doubling of the ranks, so that nb1 (the coarse bins every
    // rank scatters into) stays put: the partition kernel's scattered 16-byte stores are combined in L2, one open 128-byte line per
    // bin, and 3.7 million bins (measured on 8 GPUs with bins of constant size) thrash it: k1 49 -> 94 ms.  The gathered bins grow
    // instead; the dedup split takes them in several passes over ranges of fine ids.
    const int coarse_log2 = 6;                                                  // planned counting bins per coarse bin and rank
    int fine_bits = (W == 1) ? 9 : FINE_BITS_W2;
    if (fused) fine_bits = 5;
    if (W == 1 && !fused) for (int r = 1; r < n_ranks && fine_bits < DEV_FINE_BITS_MAX_W1; r <<= 1) fine_bits++;
    if (W == 1 && p->fine_bits > 0)
    {
        if (p->fine_bits > DEV_FINE_BITS_MAX_W1) return fail (ctx, "fine_bits must be in [1,%d]", (int)DEV_FINE_BITS_MAX_W1);
        fine_bits = p->fine_bits;
    }
    uint64_t nb1;
    if (fused)
    {   // one CTA-wide table per COARSE bin (k2_fused.cu): planned k-mer occurrences per bin = 160 % of the slots, which is a
        // half-full table when three occurrences in ten are new k-mers (30x coverage); sparser data overflows into the
        // fine-split pipeline, whose fine bins (1/32 of a coarse bin) fit the tier tables
        const uint64_t occ_per_coarse = (T * (p->bin_load_pct > 0 ? p->bin_load_pct : 160)) / 100 + 1;
        nb1 = (total_kmers + occ_per_coarse - 1) / occ_per_coarse;
    }
    else
    {
        const uint64_t occ_per_bin = (T * (p->bin_load_pct > 0 ? p->bin_load_pct : (W == 1 ? 80 : 55))) / 100 + 1;
        uint64_t nbins_fine = (total_kmers + occ_per_bin - 1) / occ_per_bin; if (nbins_fine < 1) nbins_fine = 1;
        int split_log2 = fine_bits;                                             // k >= 32: the fine ids ARE the counting bins
        if (W == 1) { split_log2 = coarse_log2; for (int r = 1; r < n_ranks; r <<= 1) split_log2++; }    // nb1 stays put, the gathered bins grow
        // several ranks: the records arriving over NVLink pass through the same L2 that combines the partition kernel's scattered
        // stores (k1 49 -> 72 ms on 2 GPUs while pieces travel): half as many open lines; the pre-split of the gathered bins
        // (count_bins_impl) takes one more leading bit
        if (W == 1 && n_ranks > 1) split_log2++;
        nb1 = (nbins_fine + (1ULL << split_log2) - 1) >> split_log2;
    }
    if (nb1 < 1) nb1 = 1;
    nb1 = (nb1 + n_ranks - 1) / n_ranks * n_ranks;                              // every rank owns nb1/n_ranks consecutive coarse bins
    if (nb1 > (1ULL << 24)) return fail (ctx, "input too large (%llu coarse bins)", (unsigned long long)nb1);
    // m-mers ranked on the device: the register scanner (k1_scan.cuh) wants w = k-m+1 a multiple of 8 and m in [8,16];
    // k < 15 falls back to the general kernel with the smallest m that still spreads the bins
    const int win = k1_fast_window (k);
    const int mg = win ? k - win + 1 : pick_device_m (k, nb1 << fine_bits);
    const int w = k - mg + 1;
    // records a rank produces for one coarse bin ~ (its k-mers * 2/(w+1)) / nb1
    double local_kmers = (double)total_kmers / n_ranks, local_reads = (double)n_reads / n_ranks;
    double est_records = local_kmers * 2.0 / (w + 1) * 1.10 + local_reads * 0.5 + 64;
    // several ranks: a (rank, bin) piece is small and the minimizer space of a multi-Gb input is crowded, so the relative
    // spread of the pieces is larger; the exchange only moves the used rounds, so head-room costs memory, not time
    // head-room: the records of one locus (~21 at 30x coverage, spread over the ranks) arrive together, so the load of a bin
    // of r records has a standard deviation near sqrt ((21/n_ranks + 1) r); 7.5 sigma covers millions of bins.  The
    // exchange only moves the used rounds, so head-room costs memory, not time; an overflow costs a second partition run.
    const double per_bin = est_records / nb1;
    uint64_t cap = (uint64_t)(per_bin + 7.5 * sqrt ((21.0 / n_ranks + 1.0) * per_bin) + per_bin * 0.10) + 64; cap = (cap + COARSE_BLK - 1) / COARSE_BLK * COARSE_BLK;
    memset (g, 0, sizeof(*g));
    g->total_kmers = total_kmers; g->nb1 = (uint32_t)nb1; g->cap = (uint32_t)cap; g->fine_bits = fine_bits; g->table_log2 = table_log2;
    g->m_device = mg; g->w = w; g->maxlen = (W == 1) ? DEV_MAXLEN_W1 : 60; g->words = W;                  // maxlen: Sequence2SuperKmer.hpp:147
    g->n_ranks = n_ranks; g->bins_per_rank = (uint32_t)(nb1 / n_ranks); g->record_bytes = 16 * W; g->coarse_blk = COARSE_BLK;
    return 0;
}

// caller-supplied geometry (staged entry points): everything the kernels index with is checked against the parameters
static int check_geometry (gatb_gpu_ctx* ctx, const gatb_gpu_params* p, const gatb_gpu_geometry* g)
{
    if (!g) return fail (ctx, "geometry is NULL");
    const int W = (p->kmer_size < 32) ? 1 : 2;
    if (g->words != W || g->record_bytes != 16u * W) return fail (ctx, "geometry: %d-word records do not match kmer_size %d", g->words, p->kmer_size);
    if (g->n_ranks < 1 || g->n_ranks > GATB_GPU_MAX_RANKS || g->bins_per_rank < 1 || (uint64_t)g->n_ranks * g->bins_per_rank != g->nb1)
        return fail (ctx, "geometry: nb1 %u != n_ranks %u x bins_per_rank %u", g->nb1, g->n_ranks, g->bins_per_rank);
    const int fmax = (W == 1) ? DEV_FINE_BITS_MAX_W1 : FINE_BITS_W2, fmin = (W == 1) ? 0 : FINE_BITS_W2;
    if (g->fine_bits < fmin || g->fine_bits > fmax) return fail (ctx, "geometry: fine_bits %d out of range", g->fine_bits);
    if (g->table_log2 < 5 || g->table_log2 > 13) return fail (ctx, "geometry: table_log2 %d out of range", g->table_log2);
    if (g->cap == 0 || g->cap % COARSE_BLK || g->coarse_blk != COARSE_BLK) return fail (ctx, "geometry: cap %u is not a multiple of %d", g->cap, (int)COARSE_BLK);
    if (g->m_device < 1 || g->m_device > 16 || g->m_device >= p->kmer_size || g->w != p->kmer_size - g->m_device + 1)
        return fail (ctx, "geometry: device minimizer %d / window %d do not match kmer_size %d", g->m_device, g->w, p->kmer_size);
    if (g->maxlen < 1 || g->maxlen > ((W == 1) ? DEV_MAXLEN_W1 : 60)) return fail (ctx, "geometry: maxlen %d out of range", g->maxlen);
    return 0;
}

// ---- stage 1: k1 into caller-provided buffers (no retry here).  h_stats: valid, invalid, stored, dropped ------------
// chunks of reads whose host->device copy is still in flight on the copy stream: k1 of chunk c waits for ready[c]
struct ReadChunks { int n; uint64_t first[17]; cudaEvent_t* ready; };

static int partition_impl (gatb_gpu_ctx* ctx, const gatb_gpu_params* p, const gatb_gpu_geometry* g,
                           const uint8_t* d_reads, const uint64_t* d_offsets, uint64_t n_reads, const uint32_t* d_nmask,
                           void* d_bins, uint32_t* d_cursors, unsigned long long* h_stats,
                           const ReadChunks* chunks = 0, uint64_t first_read = 0, uint64_t max_len = 0, const uint64_t* d_bin_off = 0)
{
    LaunchCtx L = lctx (ctx);
    if (ensure (ctx, S_STATS, 64 * 8)) return 1;
    unsigned long long* d_stats = (unsigned long long*)ctx->slot[S_STATS];
    const uint64_t nbins = (uint64_t)g->nb1 << g->fine_bits;
    K1Params k1; memset (&k1, 0, sizeof(k1));
    k1.bin_off = d_bin_off;
    k1.words = (const uint64_t*)d_reads; k1.offsets = d_offsets; k1.nmask = d_nmask; k1.n_reads = n_reads; k1.read_len = p->read_len;
    k1.k = p->kmer_size; k1.m = g->m_device; k1.w = g->w; k1.maxlen = g->maxlen;
    k1.mmask = (g->m_device >= 16) ? 0xFFFFFFFFu : ((1u << (2 * g->m_device)) - 1); k1.mask_ma1 = 0;
    k1.force_general = (p->path_flags & GATB_PATH_K1_GENERAL) ? 1 : 0;
    k1.max_len = (int)max_len; k1.no_staging = (p->path_flags & GATB_PATH_K1_STAGING) ? 0 : 1;
    k1.oriented = k1_oriented (p->kmer_size, g->m_device, g->w, p->path_flags) ? 1 : 0;
    k1.mode = K1_MODE_DEVICE; k1.nb1 = g->nb1; k1.n_regions = g->n_ranks; k1.bins_per_region = g->bins_per_rank;
    if (g->cap % COARSE_BLK) return fail (ctx, "geometry: cap %u is not a multiple of %d", g->cap, (int)COARSE_BLK); k1.fine_bits = g->fine_bits; k1.count_only = 0;
    k1.bins = d_bins; k1.cap = g->cap; k1.cursors = d_cursors; k1.stats = d_stats;
    CK (cudaMemsetAsync (d_cursors, 0, (size_t)g->nb1 * 4, ctx->stream));
    CK (cudaMemsetAsync (d_stats, 0, 4 * 8, ctx->stream));
    cudaEventRecord (ctx->kev[0], ctx->stream);
    if (chunks && chunks->n > 1)
    {
        for (int c = 0; c < chunks->n; c++)
        {
            if (chunks->ready) CK (cudaStreamWaitEvent (ctx->stream, chunks->ready[c], 0));
            k1.first_read = chunks->first[c]; k1.n_reads = chunks->first[c+1] - chunks->first[c];
            if (k1.n_reads) CK (launch_k1 (L, k1));
        }
    }
    else if (n_reads) { k1.first_read = first_read; CK (launch_k1 (L, k1)); }
    cudaEventRecord (ctx->kev[1], ctx->stream);
    CK (cudaMemcpyAsync (h_stats, d_stats, 4 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK (cudaStreamSynchronize (ctx->stream));
    return 0;
}

// ---- stage 4 (k3): partition key + ascending order of the emitted k-mers [0, n_range) (holes carry an all-ones key) ----
// in_key != NULL: the keys came with the items (routed items of a multi-GPU run); the result arrays live in context-owned slots
struct SortOut { DevResult dr; uint64_t* h_lo; uint64_t* h_hi; int32_t* h_cnt32; uint64_t* h_offs; uint64_t* h_hist; uint64_t n_big; int two_pass; int t_bits; };
static int sort_stage (gatb_gpu_ctx* ctx, const gatb_gpu_params* p, const uint64_t* u_lo, const uint64_t* u_hi, const uint32_t* u_cnt, const uint16_t* in_key,
                       uint64_t n_range, uint64_t n_items, const uint16_t* repart_host, bool to_host, SortOut* so, uint64_t n_keys_local = 0)
{
    LaunchCtx L = lctx (ctx);
    const int k = p->kmer_size, W = (k < 32) ? 1 : 2;
    // routed items carry LOCAL keys 0 .. n_keys_local-1 (the partitions this rank owns, in order): buckets and offsets are laid out
    // for those; the caller expands the offsets to all keys.  The result slot always has room for the offsets of all keys.
    const uint64_t n_keys_all = (uint64_t)p->nb_partitions * p->nb_passes;
    const uint64_t n_keys = n_keys_local ? n_keys_local : n_keys_all;
    const int histo_max = p->histo_max;
    const size_t item_bytes = 8 * W + 4;
    if (ensure (ctx, S_COUNTERS, 16 * 8)) return 1;
    if (ensure (ctx, S_HISTO, (size_t)(histo_max + 1) * 8)) return 1;
    unsigned long long* d_cnt = (unsigned long long*)ctx->slot[S_COUNTERS];
    // ---- k3: partition id + ascending order ----
    int t_bits = 0;
    // buckets cut along the distribution of canonical values (k3_range_of).  One key: 512..1024 k-mers on average.  Several keys
    // (GATB partitions): the k-mers of a partition share minimizers, their leading nucleotides are skewed key by key, and the
    // fullest buckets reach several times the average: up to 4096 items a CTA sorts them in shared memory, up to 16384 (the block
    // directory of the pooled scatter) they are sorted in place in global memory, beyond that the exact two-pass scatter takes over
    { const uint64_t avg = n_keys > 1 ? 768 : 1024;
      uint64_t per_key = n_items / n_keys + 1; while (t_bits < 2*k - 1 && t_bits < 24 && (per_key >> t_bits) > avg) t_bits++; }
    while (t_bits > 0 && (n_keys << t_bits) > (1ULL << 30)) t_bits--;
    const uint64_t n_buckets = n_keys << t_bits;
    // result arrays live in context-owned slots (valid until the next count on this context): no per-call cudaMalloc
    DevResult* dr = &so->dr; memset (so, 0, sizeof(*so));
    uint64_t n_alloc = n_items ? n_items : 1;
    if (ensure (ctx, S_SORTED, n_alloc * item_bytes + 64)) return 1;
    if (ensure (ctx, S_RESMISC, (n_keys_all + 1) * 8 + (size_t)(histo_max + 1) * 8 + 64)) return 1;
    void* d_sorted = ctx->slot[S_SORTED];
    dr->lo = d_sorted; dr->hi = (W == 2) ? (void*)((uint64_t*)d_sorted + n_alloc) : 0; dr->cnt = (void*)((uint64_t*)d_sorted + n_alloc * W);
    void* d_offs = ctx->slot[S_RESMISC]; void* d_hist = (void*)((uint64_t*)ctx->slot[S_RESMISC] + (n_keys_all + 1));
    dr->offs = d_offs; dr->histo = d_hist;

    if (ensure (ctx, S_BUCKETCNT, n_buckets * 4)) return 1;
    if (ensure (ctx, S_BUCKETOFF, (n_buckets + 1) * 8)) return 1;
    if (ensure (ctx, S_SCAN, scan_scratch_elems (n_buckets) * 8)) return 1;
    if (ensure (ctx, S_BIGLIST, n_buckets * 8)) return 1;
    if (n_keys > 1 && !in_key)
    {
        uint64_t rbytes = (1ULL << (2 * p->minimizer_size)) * 2;
        if (ensure (ctx, S_REPART, rbytes)) return 1;
        CK (cudaMemcpyAsync (ctx->slot[S_REPART], repart_host, rbytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    // the fine buffer is dead after k2b/k2c: it holds the bucket-ordered copy (pool blocks, or the exact layout)
    K3Params k3; memset (&k3, 0, sizeof(k3));
    k3.k = k; k3.m = p->minimizer_size; k3.W = W;
    k3.mmask = (1u << (2 * p->minimizer_size)) - 1; k3.mask_ma1 = gatb_mask_ma1 (p->minimizer_size);
    k3.repart = (const uint16_t*)ctx->slot[S_REPART]; k3.nb_partitions = p->nb_partitions; k3.nb_passes = p->nb_passes; k3.n_keys = (uint32_t)n_keys;
    k3.t_bits = t_bits; k3.n = n_range;
    k3.in_lo = u_lo; k3.in_hi = u_hi; k3.in_cnt = u_cnt; k3.in_key = in_key;
    k3.bucket_count = (uint32_t*)ctx->slot[S_BUCKETCNT];
    k3.bucket_off = (const uint64_t*)ctx->slot[S_BUCKETOFF];
    k3.out_lo = (uint64_t*)dr->lo; k3.out_hi = (uint64_t*)dr->hi; k3.out_cnt = (int32_t*)dr->cnt;
    k3.n_buckets = (uint32_t)n_buckets; k3.big_list = (unsigned long long*)ctx->slot[S_BIGLIST]; k3.counters = d_cnt + 8;
    cudaEventRecord (ctx->kev[6], ctx->stream);
    // ---- bucket scatter: pooled single pass (k3s) unless a bucket outgrows the block directory, then the exact two-pass path ----
    const bool no_pool = (p->path_flags & GATB_PATH_K3_NO_POOL) != 0;                          // test selector
    bool pooled = false;
    if (!no_pool && n_items)
    {
        // the directory covers buckets of up to 16384 items (those above the 4096 a CTA sorts in shared memory are sorted in place
        // in global memory, k3d); a tiny directory (test selector) forces the exact two-pass fallback
        const uint32_t dir_rounds = p->k3_dir_rounds > 0 ? (uint32_t)p->k3_dir_rounds : 4 * k3_sort_cap () / K3_BLK;
        const uint64_t pool_blocks = n_items / K3_BLK + n_buckets + 1;
        if (pool_blocks < (1ULL << 32))
        {
            if (ensure (ctx, S_FINE, pool_blocks * K3_BLK * 16 * W)) return 1;
            if (ensure (ctx, S_DIR, (size_t)dir_rounds * n_buckets * 4)) return 1;
            CK (cudaMemsetAsync (ctx->slot[S_DIR], 0xFF, (size_t)dir_rounds * n_buckets * 4, ctx->stream));
            CK (cudaMemsetAsync (d_cnt + 8, 0, 4 * 8, ctx->stream));
            CK (cudaMemsetAsync (k3.bucket_count, 0, n_buckets * 4, ctx->stream));
            k3.pool = (uint4*)ctx->slot[S_FINE]; k3.pool_blocks = (uint32_t)pool_blocks; k3.pool_ptr = (uint32_t*)(d_cnt + 10);
            k3.dir = (uint32_t*)ctx->slot[S_DIR]; k3.dir_rounds = dir_rounds; k3.ovf_flag = (uint32_t*)(d_cnt + 11);
            CK (launch_k3s_pool_scatter (L, k3));
            CK (launch_scan_u32_to_u64 (L, k3.bucket_count, (uint64_t*)ctx->slot[S_BUCKETOFF], n_buckets, (uint64_t*)ctx->slot[S_SCAN]));
            uint32_t flag = 0;
            CK (cudaMemcpyAsync (&flag, k3.ovf_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK (cudaStreamSynchronize (ctx->stream));
            pooled = (flag == 0);
            if (!pooled) { k3.pool = 0; k3.dir = 0; }
        }
    }
    if (!pooled)
    {
        if (ctx->slot_cap[S_FINE] < n_alloc * item_bytes) { if (ensure (ctx, S_FINE, n_alloc * item_bytes)) return 1; }
        k3.tmp_lo = (uint64_t*)ctx->slot[S_FINE]; k3.tmp_hi = (W == 2) ? k3.tmp_lo + n_alloc : 0; k3.tmp_cnt = (uint32_t*)(k3.tmp_lo + n_alloc * W);
        if (ensure (ctx, S_BUCKETOF, (n_range + 1) * 4)) return 1;
        k3.bucket_of = (uint32_t*)ctx->slot[S_BUCKETOF];
        CK (cudaMemsetAsync (k3.bucket_count, 0, n_buckets * 4, ctx->stream));
        CK (launch_k3a_classify (L, k3));
        CK (launch_scan_u32_to_u64 (L, k3.bucket_count, (uint64_t*)ctx->slot[S_BUCKETOFF], n_buckets, (uint64_t*)ctx->slot[S_SCAN]));
        CK (cudaMemsetAsync (k3.bucket_count, 0, n_buckets * 4, ctx->stream));
        CK (cudaMemsetAsync (d_cnt + 8, 0, 8, ctx->stream));
        CK (launch_k3b_scatter (L, k3));
    }
    // host sink: the sorted arrays leave in chunks of buckets while the next chunk is being sorted
    uint8_t* pin = 0; uint64_t* h_lo = 0; uint64_t* h_hi = 0; int32_t* h_cnt32 = 0; uint64_t* h_offs = 0; uint64_t* h_hist = 0;
    const int n_chunks = (to_host && n_items > (1u << 20)) ? 8 : 1;
    std::vector<uint64_t> chunk_bucket (n_chunks + 1), chunk_off (n_chunks + 1);
    for (int c = 0; c <= n_chunks; c++) chunk_bucket[c] = n_buckets * c / n_chunks;
    if (to_host)
    {
        const size_t hist_bytes = (size_t)(histo_max + 1) * 8, off_bytes = (n_keys + 1) * 8;
        const size_t need = off_bytes + hist_bytes + n_alloc * 8 * W + n_alloc * 4 + 256;
        pin = (uint8_t*) pinned_ensure (ctx, need);
        if (!pin) return fail (ctx, "pinned host allocation of the result (%zu bytes) failed", need);
        h_lo = (uint64_t*)pin; pin += n_alloc * 8;
        if (W == 2) { h_hi = (uint64_t*)pin; pin += n_alloc * 8; }
        h_offs = (uint64_t*)pin; pin += off_bytes; h_hist = (uint64_t*)pin; pin += hist_bytes; h_cnt32 = (int32_t*)pin;
        for (int c = 0; c <= n_chunks; c++)
            CK (cudaMemcpyAsync (&chunk_off[c], (const uint64_t*)ctx->slot[S_BUCKETOFF] + chunk_bucket[c], 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK (cudaStreamSynchronize (ctx->stream));
    }
    unsigned long long n_big = 0, n_big_done = 0;
    for (int c = 0; c < n_chunks; c++)
    {
        k3.bucket_begin = (uint32_t)chunk_bucket[c]; k3.bucket_end = (uint32_t)chunk_bucket[c+1];
        CK (launch_k3c_sort (L, k3));
        if (to_host || c == n_chunks - 1)
        {   // buckets of this chunk too large for the shared-memory sort: sorted in place now, so that the chunk leaves complete
            CK (cudaMemcpyAsync (&n_big, d_cnt + 8, 8, cudaMemcpyDeviceToHost, ctx->stream));
            CK (cudaStreamSynchronize (ctx->stream));
            if (n_big > n_big_done)
            {
                K3Params k3d = k3; k3d.big_list = k3.big_list + n_big_done;
                CK (launch_k3d_sort_big (L, k3d, (uint32_t)(n_big - n_big_done)));
                n_big_done = n_big;
            }
        }
        if (to_host)
        {
            CK (cudaEventRecord (ctx->cev[20 + c], ctx->stream));
            CK (cudaStreamWaitEvent (ctx->copy_stream, ctx->cev[20 + c], 0));
            const uint64_t a0 = chunk_off[c], cnt = chunk_off[c+1] - chunk_off[c];
            if (cnt)
            {
                CK (cudaMemcpyAsync (h_lo + a0, (const uint64_t*)dr->lo + a0, cnt * 8, cudaMemcpyDeviceToHost, ctx->copy_stream));
                if (W == 2) CK (cudaMemcpyAsync (h_hi + a0, (const uint64_t*)dr->hi + a0, cnt * 8, cudaMemcpyDeviceToHost, ctx->copy_stream));
                CK (cudaMemcpyAsync (h_cnt32 + a0, (const int32_t*)dr->cnt + a0, cnt * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
            }
        }
    }
    cudaEventRecord (ctx->kev[7], ctx->stream);
    // part_offsets[key] = bucket_off[key << t_bits]
    {
        std::vector<uint64_t> offs (n_keys + 1);
        // one strided copy: entry key << t_bits of the bucket offsets for every key (and the total at the end)
        CK (cudaMemcpy2DAsync (offs.data (), 8, ctx->slot[S_BUCKETOFF], (size_t)8 << t_bits, 8, n_keys + 1, cudaMemcpyDeviceToHost, ctx->stream));
        CK (cudaStreamSynchronize (ctx->stream));
        CK (cudaMemcpyAsync (d_offs, offs.data (), (n_keys + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK (cudaMemcpyAsync (d_hist, ctx->slot[S_HISTO], (size_t)(histo_max + 1) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        if (to_host)
        {
            memcpy (h_offs, offs.data (), (n_keys + 1) * 8);
            CK (cudaMemcpyAsync (h_hist, ctx->slot[S_HISTO], (size_t)(histo_max + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CK (cudaStreamSynchronize (ctx->copy_stream));
        }
    }
    so->h_lo = h_lo; so->h_hi = h_hi; so->h_cnt32 = h_cnt32; so->h_offs = h_offs; so->h_hist = h_hist;
    so->n_big = n_big; so->two_pass = pooled ? 0 : 1; so->t_bits = t_bits;
    return 0;
}

// ---- stages 2-4: fine split of nb1_local coarse bins gathered from n_src sources, count, partition id + sort --------
static int count_bins_impl (gatb_gpu_ctx* ctx, const gatb_gpu_params* p, const gatb_gpu_geometry* g, int n_src,
                            const void* const* d_src_bins, const uint32_t* const* d_src_cursors,
                            uint32_t nb1_local, const uint16_t* repart_host, uint64_t total_kmers_bound, gatb_gpu_result* out,
                            bool to_host = false, const uint64_t* const* d_src_off = 0, int route_ranks = 0, uint64_t* route_counts = 0, uint16_t** route_keys = 0)
{
    LaunchCtx L = lctx (ctx);
    const int k = p->kmer_size, W = g->words;
    const uint64_t n_keys = (uint64_t)p->nb_partitions * p->nb_passes;
    const int histo_max = p->histo_max, fine_bits = g->fine_bits, table_log2 = g->table_log2;
    const uint64_t nbins = (uint64_t)nb1_local << fine_bits;
    const uint32_t cap = d_src_off ? 0xFFFFFFFFu : g->cap;        // dense sources hold exactly what their cursors say
    const size_t rec_bytes = 16 * W;
    if (n_src < 1 || n_src > GATB_GPU_MAX_SOURCES) return fail (ctx, "n_src must be in [1,%d]", GATB_GPU_MAX_SOURCES);

    // ---- compact layout of the fine-split copy: coarse_off = exclusive scan of the gathered record counts ----
    if (ensure (ctx, S_TOTCUR, (size_t)nb1_local * 4)) return 1;
    if (ensure (ctx, S_COARSEOFF, ((size_t)nb1_local + 1) * 8)) return 1;
    if (ensure (ctx, S_SCAN, scan_scratch_elems (nb1_local) * 8)) return 1;          // (the sort stage sizes it again for its buckets)
    const uint64_t desc_cap = nbins + nbins / 8 + 1024;          // (adaptive bins: at most one per fine id, plus the bins of the two-pass fallback)
    if (ensure (ctx, S_BINDESC, desc_cap * 8)) return 1;
    CursorList CL; CL.n = n_src; for (int s = 0; s < n_src; s++) CL.cur[s] = d_src_cursors[s];
    if (ensure (ctx, S_COUNTERS, 16 * 8)) return 1;
    uint32_t* d_maxbin = (uint32_t*)((unsigned long long*)ctx->slot[S_COUNTERS] + 15);
    CK (cudaMemsetAsync (d_maxbin, 0, 8, ctx->stream));
    k_sum_cursors<<<(nb1_local + 255) / 256, 256, 0, ctx->stream>>> (CL, nb1_local, cap, (uint32_t*)ctx->slot[S_TOTCUR], d_maxbin); ctx->launches++;
    CK (launch_scan_u32_to_u64 (L, (const uint32_t*)ctx->slot[S_TOTCUR], (uint64_t*)ctx->slot[S_COARSEOFF], nb1_local, (uint64_t*)ctx->slot[S_SCAN]));
    uint64_t n_records = 0; uint32_t max_bin = 0;
    CK (cudaMemcpyAsync (&n_records, (const uint64_t*)ctx->slot[S_COARSEOFF] + nb1_local, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK (cudaMemcpyAsync (&max_bin, d_maxbin, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK (cudaStreamSynchronize (ctx->stream));
    cudaEventRecord (ctx->ev[2], ctx->stream);

    // ---- k2a: fine split (not on the fused path: k <= 31 counts straight out of the coarse bins) ----
    const bool fused = path_fused (p);
    const int dedup = (p->path_flags & GATB_PATH_NO_DEDUP) ? 0 : 1;
    if (!fused && ensure (ctx, S_FINE, (n_records + 1) * rec_bytes)) return 1;
    K2aSrc S2; memset (&S2, 0, sizeof(S2)); S2.n = n_src;
    for (int s = 0; s < n_src; s++) { S2.bins[s] = (const uint4*)d_src_bins[s]; S2.cursors[s] = d_src_cursors[s]; S2.off[s] = d_src_off ? d_src_off[s] : 0; }
    cudaEventRecord (ctx->kev[2], ctx->stream);
    uint64_t n_unique_records = n_records;
    bool presplit_used = false;
    uint64_t nbins_count = nbins;                  // bins the counting kernels see (the dedup split forms its own: adaptive bins)
    bool desc_abs = false;                         // their descriptors hold absolute record offsets
    if (fused) cudaEventRecord (ctx->kev[3], ctx->stream);
    else if (W == 1 && dedup && n_records)
    {   // ---- k <= 31: the bin is staged in shared memory once, identical records collapse, multiplicities travel with the records ----
        if (n_records >= (1ULL << 32)) return fail (ctx, "too many super-k-mer records for one device (%llu): use more passes or more GPUs", (unsigned long long)n_records);
        if (ensure (ctx, S_OVFLIST, (nbins > nb1_local ? nbins : nb1_local) * 4)) return 1;
        unsigned long long* d_k2a = (unsigned long long*)ctx->slot[S_COUNTERS] + 12;
        CK (cudaMemsetAsync (d_k2a, 0, 3 * 8, ctx->stream));
        // ---- bins gathered from several ranks are several times larger than a CTA's staging area: instead of taking them in as many
        //      passes (each reads the whole bin again: 66 ms for 5 passes on 8 GPUs), a plain two-pass split by the leading bits of the
        //      fine id first cuts them into sub-bins of single-GPU size (one more copy of the records, ~18 ms), which the dedup split
        //      then takes from that dense copy in one pass each ----
        uint32_t nb_d = nb1_local, cap_d = cap, max_bin_d = max_bin; int fine_bits_d = fine_bits;
        const uint64_t* coarse_off_d = (const uint64_t*)ctx->slot[S_COARSEOFF];
        K2aSrc S2d = S2;
        {
            const uint32_t mean_bin = (uint32_t)(n_records / (nb1_local ? nb1_local : 1)) + 1;
            int sub_bits = 0;
            if (n_src > 1 || (p->path_flags & GATB_PATH_K2A_PRESPLIT))
                while (sub_bits < fine_bits - 6 && ((mean_bin >> sub_bits) * 5) / 4 > k2a_two_cta_capacity (fine_bits - sub_bits)) sub_bits++;
            if ((p->path_flags & GATB_PATH_K2A_PRESPLIT) && sub_bits == 0 && fine_bits > 2) sub_bits = 2;        // test selector
            if (sub_bits)
            {
                const uint64_t nsub = (uint64_t)nb1_local << sub_bits;
                if (ensure (ctx, S_PRESPLIT, (n_records + 1) * rec_bytes)) return 1;
                if (ensure (ctx, S_SUBOFF, (nsub + 1) * 8)) return 1;
                if (ensure (ctx, S_SUBCNT, nsub * 4)) return 1;
                K2aPresplit PS; PS.sub_off = (uint64_t*)ctx->slot[S_SUBOFF]; PS.sub_cnt = (uint32_t*)ctx->slot[S_SUBCNT]; PS.shift = fine_bits - sub_bits;
                CK (launch_k2a_split (L, W, S2, ctx->slot[S_PRESPLIT], (const uint64_t*)ctx->slot[S_COARSEOFF], nb1_local, cap, sub_bits, 0, 0, 0, 0, 0, &PS));
                CursorList C1; C1.n = 1; C1.cur[0] = PS.sub_cnt;
                CK (cudaMemsetAsync (d_maxbin, 0, 8, ctx->stream));
                if (ensure (ctx, S_TOTCUR, nsub * 4)) return 1;
                k_sum_cursors<<<(unsigned)((nsub + 255) / 256), 256, 0, ctx->stream>>> (C1, (uint32_t)nsub, 0xFFFFFFFFu, (uint32_t*)ctx->slot[S_TOTCUR], d_maxbin); ctx->launches++;
                CK (cudaMemcpyAsync (&max_bin_d, d_maxbin, 4, cudaMemcpyDeviceToHost, ctx->stream));
                CK (cudaStreamSynchronize (ctx->stream));
                memset (&S2d, 0, sizeof(S2d)); S2d.n = 1; S2d.bins[0] = (const uint4*)ctx->slot[S_PRESPLIT]; S2d.cursors[0] = PS.sub_cnt; S2d.off[0] = PS.sub_off;
                presplit_used = true;
                nb_d = (uint32_t)nsub; cap_d = 0xFFFFFFFFu; fine_bits_d = fine_bits - sub_bits; coarse_off_d = PS.sub_off;
            }
        }
        const uint32_t rmax = (p->path_flags & GATB_PATH_K2A_SMALL_STAGE) ? 256u : k2a_dedup_rmax (max_bin_d, (uint32_t)(n_records / (nb_d ? nb_d : 1)) + 1, fine_bits_d);
        // counting bins of <= target k-mers of surviving records (+ one fine id's worth): the distinct k-mers of a bin (at most
        // that many) must stay below 3/4 of the table, and a fuller table probes longer: 45 % measured best
        // (k2b 62.2 + 1.1 ms for the 0.2 % of bins that still overflow; 60 %: 65.4 + 3.8, 30 %: 63.0 + 0.4)
        const uint32_t target = (uint32_t)((((uint64_t)1 << table_log2) * (p->bin_target_pct > 0 ? p->bin_target_pct : 45)) / 100);
        // the distinct k-mers of a bin are at most the k-mers of its records, typically half of them (the variants of a locus' record --
        // reads that end inside it, sequencing errors -- repeat most of its k-mers): a bin whose records hold more than 0.95 T k-mers
        // (one fine id = one minimizer value holds several loci of a multi-Gb genome) rarely fits the 3/4 T slots a table may fill: measured on the 8-GPU geometry, a threshold of 0.95 T costs 38 + 43 ms (first tier + tiers), 1.4 T 54 + 36 ms -- a failed attempt in the first tier is dearer than a needless trip to the 1024-slot tier
        const uint32_t big_load = (k2b_variant (p->path_flags) == 1) ? (uint32_t)((((uint64_t)1 << table_log2) * 95) / 100) : 0u;
        CK (launch_k2a_dedup_split (L, S2d, ctx->slot[S_FINE], coarse_off_d, nb_d, cap_d, fine_bits_d,
                                    (uint2*)ctx->slot[S_BINDESC], rmax, (uint32_t*)ctx->slot[S_OVFLIST], d_k2a, target, big_load));
        unsigned long long h_k2a[3] = { 0, 0, 0 };
        CK (cudaMemcpyAsync (h_k2a, d_k2a, 24, cudaMemcpyDeviceToHost, ctx->stream));
        CK (cudaStreamSynchronize (ctx->stream));
        if (h_k2a[0])
        {   // bins the kernel gave up on (a skewed range of fine ids, more passes than it tracks) go through the plain two-pass
            // split (multiplicity 1, one counting bin per fine id), their descriptors appended after the adaptive ones
            if (h_k2a[2] + (h_k2a[0] << fine_bits_d) > desc_cap) return fail (ctx, "bin descriptors exhausted");
            CK (launch_k2a_split (L, W, S2d, ctx->slot[S_FINE], coarse_off_d, nb_d, cap_d, fine_bits_d,
                                  (uint2*)ctx->slot[S_BINDESC], (const uint32_t*)ctx->slot[S_OVFLIST], (uint32_t)h_k2a[0], 1, h_k2a[2]));
        }
        cudaEventRecord (ctx->kev[3], ctx->stream);
        n_unique_records = h_k2a[1];                 // (bins of the two-pass kernel are not in this figure)
        nbins_count = h_k2a[2] + (h_k2a[0] << fine_bits_d);
        desc_abs = true;
    }
    else
    {
        CK (launch_k2a_split (L, W, S2, ctx->slot[S_FINE], (const uint64_t*)ctx->slot[S_COARSEOFF], nb1_local, cap, fine_bits, (uint2*)ctx->slot[S_BINDESC]));
        cudaEventRecord (ctx->kev[3], ctx->stream);
    }
    cudaEventRecord (ctx->ev[3], ctx->stream);

    // ---- k2b: count.  When the single source is the context's own coarse buffer it is dead now and becomes the output. ----
    const uint32_t amin = p->abundance_min < 1 ? 1 : (uint32_t)p->abundance_min;
    const uint32_t amax = p->abundance_max < 0 ? 0x7fffffffu : (uint32_t)p->abundance_max;
    const uint32_t emin = p->emit_all ? 1u : amin, emax = p->emit_all ? 0xffffffffu : amax;
    // (the fused kernel reads the coarse bins while it emits; after a pre-split the dense copy is dead once the dedup split has run)
    const int S_OUT = (!fused && n_src == 1 && d_src_bins[0] == ctx->slot[S_COARSE]) ? S_COARSE : (presplit_used ? S_PRESPLIT : S_UNSORTED);
    uint64_t out_bound = total_kmers_bound / emin + 1;                         // a k-mer emitted needs >= emin occurrences
    const size_t item_bytes = 8 * W + 4;
    const uint64_t block_slack = (uint64_t)(ctx->sm_count * 8 + 8) * 8 * 2048;  // every warp of k2b reserves output in blocks of 2048 slots
    uint64_t out_cap = total_kmers_bound / 8 + 4096;                           // first guess; the retry below corrects it
    if (out_cap > out_bound) out_cap = out_bound;
    out_cap += block_slack;
    if (out_cap < ctx->slot_cap[S_OUT] / item_bytes) out_cap = ctx->slot_cap[S_OUT] / item_bytes;   // never shrink: no re-allocation per call

    if (ensure (ctx, S_HISTO, (size_t)(histo_max + 1) * 8)) return 1;
    { const uint64_t nl = nbins_count > nbins ? nbins_count : nbins;
      if (ensure (ctx, S_OVFLIST, (nl > nb1_local ? nl : nb1_local) * 4)) return 1;
      if (ensure (ctx, S_OVFLIST2, nl * 4)) return 1; }
    unsigned long long* d_cnt = (unsigned long long*)ctx->slot[S_COUNTERS];
    unsigned long long h_cnt[16];
    uint64_t n_ovf = 0, n_ovf_first = 0, tier_out[3] = { 0, 0, 0 };          // bins leaving the warp tier / the two CTA tiers
    uint64_t* u_lo = 0; uint64_t* u_hi = 0; uint32_t* u_cnt = 0;
    K2Params k2;
    for (int attempt = 0; ; attempt++)
    {
        if (ensure (ctx, S_OUT, out_cap * item_bytes)) return 1;
        u_lo = (uint64_t*)ctx->slot[S_OUT];
        u_hi = (W == 2) ? u_lo + out_cap : 0;
        u_cnt = (uint32_t*)(u_lo + out_cap * W);
        CK (cudaMemsetAsync (ctx->slot[S_HISTO], 0, (size_t)(histo_max + 1) * 8, ctx->stream));
        CK (cudaMemsetAsync (d_cnt, 0, 16 * 8, ctx->stream));
        memset (&k2, 0, sizeof(k2));
        k2.k = k; k2.W = W; k2.recs = ctx->slot[S_FINE]; k2.bin_desc = (const uint2*)ctx->slot[S_BINDESC]; k2.nbins = (uint32_t)nbins_count;
        // (absolute descriptors: every bin takes coarse_off[bin >> 31] = coarse_off[0] = 0 as its base)
        k2.coarse_off = (const uint64_t*)ctx->slot[S_COARSEOFF]; k2.fine_bits = desc_abs ? 31 : fine_bits; k2.table_log2 = table_log2;
        k2.path_flags = p->path_flags; k2.oriented = k1_oriented (k, g->m_device, g->w, p->path_flags) ? 1 : 0;
        k2.emit_min = emin; k2.emit_max = emax; k2.solid_min = amin; k2.solid_max = amax; k2.histo_max = histo_max;
        k2.histogram = (unsigned long long*)ctx->slot[S_HISTO];
        k2.out_lo = u_lo; k2.out_hi = u_hi; k2.out_cnt = u_cnt; k2.out_cap = out_cap;
        k2.counters = d_cnt; k2.ovf_list = (uint32_t*)ctx->slot[S_OVFLIST]; k2.ovf_counter = 4;
        cudaEventRecord (ctx->kev[4], ctx->stream);
        if (fused) CK (launch_k2f_count (L, k2, S2, nb1_local, cap, nb1_local, dedup));
        else       CK (launch_k2b_count (L, k2));
        cudaEventRecord (ctx->kev[5], ctx->stream);
        CK (cudaMemcpyAsync (h_cnt, d_cnt, 8 * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK (cudaStreamSynchronize (ctx->stream));
        n_ovf = h_cnt[4];
        n_ovf_first = n_ovf;
        if (n_ovf) cudaEventRecord (ctx->kev[8], ctx->stream);
        int cur_list = S_OVFLIST;
        if (fused && n_ovf)
        {   // ---- coarse bins whose distinct k-mers outgrew the CTA's table: fine split of just those bins (multiplicity 1), then
            //      their fine bins (1 << fine_bits each) go through the tier kernels below ----
            if (ensure (ctx, S_FINE, (n_records + 1) * rec_bytes)) return 1;
            CK (launch_k2a_split (L, W, S2, ctx->slot[S_FINE], (const uint64_t*)ctx->slot[S_COARSEOFF], nb1_local, cap, fine_bits,
                                  (uint2*)ctx->slot[S_BINDESC], (const uint32_t*)ctx->slot[S_OVFLIST], (uint32_t)n_ovf));
            k_expand_bins<<<(unsigned)(((n_ovf << fine_bits) + 255) / 256), 256, 0, ctx->stream>>> ((const uint32_t*)ctx->slot[S_OVFLIST], (uint32_t)n_ovf, fine_bits,
                                                                                                      (uint32_t*)ctx->slot[S_OVFLIST2]); ctx->launches++;
            k2.recs = ctx->slot[S_FINE];
            n_ovf <<= fine_bits; cur_list = S_OVFLIST2;
            k2.ovf_list = (uint32_t*)ctx->slot[cur_list];
        }
        const bool no_tier2 = (p->path_flags & GATB_PATH_NO_TIER2) != 0;                      // test selector: straight to the global table
        if (n_ovf && ((W == 1 && (fused || k2b_variant (p->path_flags) == 1)) || W == 2) && !no_tier2)
        {   // ---- further tiers: the bins a warp's 2^table_log2-slot table could not hold are counted by CTAs with 2048,
            //      then 8192 slots (k2b_count_w1 over a bin list); what still overflows goes to the global table ----
            // (the claimed-slot lists of k2b_count_w1 are per warp, 3/32 of the table each: 4096 slots leave a warp 384 of them)
            // (first, warps again with 2048-slot tables -- k2b_warp_bins over the bin list, at the rate of the first tier: on multi-Gb
            //  inputs one bin in eleven overflows 512 slots, and CTA tiers took 69 ms for them on 8 GPUs)
            const int tier_log2[3] = { 10, 12, 13 }, tier_counter[3] = { 13, 7, 12 };
            for (int t = 0; t < 3 && n_ovf; t++)
            {
                if (!fused && tier_log2[t] <= table_log2) continue;
                if ((fused || W == 2) && tier_log2[t] < 12) continue;
                const int other = (cur_list == S_OVFLIST) ? S_OVFLIST2 : S_OVFLIST;
                K2Params k2t = k2;
                k2t.bin_list = (const uint32_t*)ctx->slot[cur_list]; k2t.n_list = (uint32_t)n_ovf;
                k2t.ovf_list = (uint32_t*)ctx->slot[other]; k2t.ovf_counter = tier_counter[t]; k2t.table_log2 = tier_log2[t];
                CK (cudaMemsetAsync (d_cnt + 3, 0, 8, ctx->stream));                 // work counter
                CK (launch_k2b_count_list (L, k2t));
                CK (cudaMemcpyAsync (h_cnt, d_cnt, 16 * 8, cudaMemcpyDeviceToHost, ctx->stream));
                CK (cudaStreamSynchronize (ctx->stream));
                n_ovf = h_cnt[tier_counter[t]];
                tier_out[t] = n_ovf;
                cur_list = other;
            }
            k2.ovf_list = (uint32_t*)ctx->slot[cur_list];                            // what is left goes to the global table
            if (!n_ovf) cudaEventRecord (ctx->kev[9], ctx->stream);
        }
        if (n_ovf)
        {   // ---- k2c: bins that did not fit the shared-memory table share one global table ----
            CK (launch_k2c_measure (L, k2, (uint32_t)n_ovf));
            CK (cudaMemcpyAsync (h_cnt, d_cnt, 8 * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CK (cudaStreamSynchronize (ctx->stream));
            uint64_t occ = h_cnt[5]; int g_log2 = 10; while ((1ULL << g_log2) < 2 * occ) g_log2++;
            if (g_log2 > 31) return fail (ctx, "fallback table too large (%llu k-mers in overflowing bins)", (unsigned long long)occ);
            size_t gT = (size_t)1 << g_log2;
            if (ensure (ctx, S_GTABLE, gT * (8 * W + 4))) return 1;
            k2.g_lo = (uint64_t*)ctx->slot[S_GTABLE]; k2.g_hi = (W == 2) ? k2.g_lo + gT : 0; k2.g_cnt = (uint32_t*)(k2.g_lo + gT * W); k2.g_log2 = g_log2;
            if (W == 1) CK (cudaMemsetAsync (k2.g_lo, 0xFF, gT * 8, ctx->stream));
            else      { CK (cudaMemsetAsync (k2.g_lo, 0, gT * 8, ctx->stream)); CK (cudaMemsetAsync (k2.g_hi, 0xFF, gT * 8, ctx->stream)); }
            CK (cudaMemsetAsync (k2.g_cnt, 0, gT * 4, ctx->stream));
            CK (launch_k2c_insert (L, k2, (uint32_t)n_ovf));
            CK (launch_k2c_scan (L, k2));
            cudaEventRecord (ctx->kev[9], ctx->stream);
            CK (cudaMemcpyAsync (h_cnt, d_cnt, 8 * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CK (cudaStreamSynchronize (ctx->stream));
        }
        if (h_cnt[0] <= out_cap) break;
        if (attempt >= 1) return fail (ctx, "output capacity exceeded: %llu k-mers to emit, room for %llu", (unsigned long long)h_cnt[0], (unsigned long long)out_cap);
        out_cap = h_cnt[0] + block_slack;                     // the cursor kept counting: it bounds the demand -> count again
    }
    const uint64_t n_items = h_cnt[6];          // k-mers emitted
    const uint64_t n_range = h_cnt[0];          // extent of the unsorted array (block reservations leave EMPTY holes)
    cudaEventRecord (ctx->ev[4], ctx->stream);

    SortOut so; memset (&so, 0, sizeof(so));
    if (route_ranks > 0)
    {   // ---- several GPUs, second exchange: instead of sorting here, the emitted k-mers are grouped by the rank that owns their
        //      partition (key % route_ranks) together with their keys; the caller exchanges the groups and every rank sorts the
        //      partitions it owns (gatb_gpu_sort_routed) ----
        if (route_ranks > GATB_GPU_MAX_RANKS) return fail (ctx, "route_ranks must be in [1,%d]", GATB_GPU_MAX_RANKS);
        if (ensure (ctx, S_ROUTECNT, 136 * 8)) return 1;                      // cursor of destination r at [16 r] (one 128-byte line each), flag at [128]
        unsigned long long* d_rc = (unsigned long long*)ctx->slot[S_ROUTECNT];
        K3Params k3; memset (&k3, 0, sizeof(k3));
        k3.k = k; k3.m = p->minimizer_size; k3.W = W;
        k3.mmask = (1u << (2 * p->minimizer_size)) - 1; k3.mask_ma1 = gatb_mask_ma1 (p->minimizer_size);
        k3.nb_partitions = p->nb_partitions; k3.nb_passes = p->nb_passes; k3.n_keys = (uint32_t)n_keys; k3.n = n_range;
        k3.in_lo = u_lo; k3.in_hi = u_hi; k3.in_cnt = u_cnt;
        if (n_keys > 1)
        {
            uint64_t rbytes = (1ULL << (2 * p->minimizer_size)) * 2;
            if (ensure (ctx, S_REPART, rbytes)) return 1;
            CK (cudaMemcpyAsync (ctx->slot[S_REPART], repart_host, rbytes, cudaMemcpyHostToDevice, ctx->stream));
            k3.repart = (const uint16_t*)ctx->slot[S_REPART];
        }
        cudaEventRecord (ctx->kev[6], ctx->stream);
        // regions of dest_cap items per destination: a quarter of head-room over the even share; if one fills up, once more with
        // regions that hold everything
        uint64_t dest_cap = n_items / route_ranks + n_items / (4 * (uint64_t)route_ranks) + 4096;
        if (dest_cap > n_items + 1) dest_cap = n_items + 1;
        uint64_t* o_lo = 0; uint64_t* o_hi = 0; uint32_t* o_cnt = 0; uint16_t* o_key = 0;
        unsigned long long h_rc[9];
        for (int attempt = 0; ; attempt++)
        {
            const uint64_t n_alloc = dest_cap * route_ranks;
            // (the records k2b counted are dead by now: their buffer takes the routed items; the sort of the received items takes it next)
            if (ensure (ctx, S_FINE, n_alloc * (8 * W + 6) + 64)) return 1;
            o_lo = (uint64_t*)ctx->slot[S_FINE]; o_hi = (W == 2) ? o_lo + n_alloc : 0;
            o_cnt = (uint32_t*)(o_lo + n_alloc * W); o_key = (uint16_t*)(o_cnt + n_alloc);
            CK (cudaMemsetAsync (d_rc, 0, 136 * 8, ctx->stream));
            CK (launch_k3r_route (L, k3, (uint32_t)route_ranks, dest_cap, d_rc, o_lo, o_hi, o_cnt, o_key, (uint32_t*)(d_rc + 128)));
            unsigned long long h_all[129];
            CK (cudaMemcpyAsync (h_all, d_rc, 129 * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CK (cudaStreamSynchronize (ctx->stream));
            for (int r = 0; r < 8; r++) h_rc[r] = h_all[16 * r];
            h_rc[8] = h_all[128];
            if (!(h_rc[8] & 1)) break;
            if (attempt >= 1) return fail (ctx, "routing: a destination region overflowed twice");
            dest_cap = n_items + 1;
        }
        unsigned long long run = 0;
        for (int r = 0; r < route_ranks; r++) { run += h_rc[r]; if (route_counts) route_counts[r] = h_rc[r]; }
        if (run != n_items) return fail (ctx, "routing: %llu items counted, %llu emitted", run, (unsigned long long)n_items);
        if (route_counts) route_counts[route_ranks] = dest_cap;
        cudaEventRecord (ctx->kev[7], ctx->stream);
        so.dr.lo = o_lo; so.dr.hi = o_hi; so.dr.cnt = o_cnt; so.dr.offs = 0; so.dr.histo = ctx->slot[S_HISTO];
        if (route_keys) *route_keys = o_key;
    }
    else if (sort_stage (ctx, p, u_lo, u_hi, u_cnt, 0, n_range, n_items, repart_host, to_host, &so)) return 1;
    DevResult* dr = &so.dr;
    uint64_t* h_lo = so.h_lo; uint64_t* h_hi = so.h_hi; int32_t* h_cnt32 = so.h_cnt32; uint64_t* h_offs = so.h_offs; uint64_t* h_hist = so.h_hist;
    cudaEventRecord (ctx->ev[5], ctx->stream);
    CK (cudaStreamSynchronize (ctx->stream));

    // ---- result ----
    memset (out, 0, sizeof(*out));
    out->n_keys = n_keys; out->n_items = n_items; out->on_device = 1; out->owner = 0;
    out->part_offsets = (uint64_t*)dr->offs; out->kmers_lo = (uint64_t*)dr->lo; out->kmers_hi = (uint64_t*)dr->hi;
    out->counts = (int32_t*)dr->cnt; out->histogram = (uint64_t*)dr->histo;
    if (to_host)
    {
        out->on_device = 0;
        out->part_offsets = h_offs; out->kmers_lo = h_lo; out->kmers_hi = h_hi; out->counts = h_cnt32; out->histogram = h_hist;
    }
    out->stats[GATB_STAT_DISTINCT] = h_cnt[1]; out->stats[GATB_STAT_SOLID] = h_cnt[2];
    out->stats[GATB_STAT_RECORDS] = n_records; out->stats[GATB_STAT_BINS] = nbins_count; out->stats[GATB_STAT_OVERFLOW_BINS] = n_ovf_first; out->stats[12] = n_ovf;
    out->stats[14] = tier_out[0]; out->stats[15] = tier_out[1];
    out->stats[GATB_STAT_RECORD_BYTES] = n_records * rec_bytes; out->stats[GATB_STAT_UNIQUE_RECORDS] = n_unique_records;
    float ms;
    cudaEventElapsedTime (&ms, ctx->ev[2], ctx->ev[3]); out->seconds[2] = ms * 1e-3;
    cudaEventElapsedTime (&ms, ctx->ev[3], ctx->ev[4]); out->seconds[3] = ms * 1e-3;
    cudaEventElapsedTime (&ms, ctx->ev[4], ctx->ev[5]); out->seconds[4] = ms * 1e-3;
    for (int i = 1; i < 4; i++) { cudaEventElapsedTime (&ms, ctx->kev[2*i], ctx->kev[2*i+1]); out->kernel_seconds[i] = ms * 1e-3; }
    if (n_ovf_first) { cudaEventElapsedTime (&ms, ctx->kev[8], ctx->kev[9]); out->kernel_seconds[4] = ms * 1e-3; out->stats[11] = h_cnt[5]; }
    out->kernel_seconds[5] = (double)so.n_big; out->kernel_seconds[6] = so.two_pass; out->kernel_seconds[7] = so.t_bits;          // diagnostics of the sort stage
    return 0;
}

static int count_dev_impl (gatb_gpu_ctx* ctx, const gatb_gpu_params* p, const uint16_t* repart_host,
                           const uint8_t* d_reads, const uint64_t* d_offsets, uint64_t n_reads, const uint32_t* d_nmask,
                           gatb_gpu_result* out, const ReadChunks* chunks = 0, bool to_host = false)
{
    cudaEventRecord (ctx->ev[1], ctx->stream);
    uint64_t total_kmers = 0, total_nt = 0, max_len = 0;
    if (workload_size (ctx, p, d_offsets, n_reads, &total_kmers, &total_nt, &max_len)) return 1;
    if (max_len >= (1ULL << 21)) return fail (ctx, "reads longer than 2^21-1 nucleotides are not supported by the partition kernel yet (longest: %llu)", (unsigned long long)max_len);
    gatb_gpu_params p_long;
    if (max_len >= (1ULL << 20) && !(p->path_flags & GATB_PATH_K1_GENERAL))
    {   // the oriented register scanner keeps 20 bits of k-mer index per event: longer reads take the general kernel
        p_long = *p; p_long.path_flags |= GATB_PATH_K1_GENERAL; p = &p_long;
    }
    gatb_gpu_geometry g;
    if (plan_geometry (ctx, p, total_kmers, n_reads, 1, &g)) return 1;
    const uint64_t nbins = (uint64_t)g.nb1 << g.fine_bits;
    if (ensure (ctx, S_CURSORS, (size_t)g.nb1 * 4)) return 1;
    uint64_t retries = 0;
    unsigned long long h_stats[4];
    const uint64_t* d_bin_off = 0;
    {
        // first attempt: bins of the planned fixed capacity (round-interleaved layout).  If any bin overflows -- skewed inputs: a
        // metagenome whose dominant species is covered thousands of times puts 15x the average into some bins -- the cursors hold
        // the exact demand of every bin: the second run lays the bins out DENSELY at their exact sizes (no capacity to guess, no
        // memory wasted on the largest bin), and the readers follow the offset table.
        const size_t planned = (size_t)g.nb1 * g.cap * g.record_bytes;
        if (planned > ((size_t)150 << 30)) return fail (ctx, "partition buffer of %zu bytes does not fit the device", planned);
        if (ensure (ctx, S_COARSE, planned)) return 1;
        if (partition_impl (ctx, p, &g, d_reads, d_offsets, n_reads, d_nmask, ctx->slot[S_COARSE], (uint32_t*)ctx->slot[S_CURSORS],
                            h_stats, chunks, 0, max_len)) return 1;
        if (h_stats[3] != 0)
        {
            retries = 1;
            if (ensure (ctx, S_BINOFF, ((size_t)g.nb1 + 1) * 8)) return 1;
            if (ensure (ctx, S_SCAN, scan_scratch_elems (g.nb1) * 8)) return 1;
            CK (launch_scan_u32_to_u64 (lctx (ctx), (const uint32_t*)ctx->slot[S_CURSORS], (uint64_t*)ctx->slot[S_BINOFF], g.nb1, (uint64_t*)ctx->slot[S_SCAN]));
            uint64_t total_records = 0;
            CK (cudaMemcpyAsync (&total_records, (const uint64_t*)ctx->slot[S_BINOFF] + g.nb1, 8, cudaMemcpyDeviceToHost, ctx->stream));
            CK (cudaStreamSynchronize (ctx->stream));
            if (ensure (ctx, S_COARSE, (total_records + 1) * g.record_bytes)) return 1;
            d_bin_off = (const uint64_t*)ctx->slot[S_BINOFF];
            if (partition_impl (ctx, p, &g, d_reads, d_offsets, n_reads, d_nmask, ctx->slot[S_COARSE], (uint32_t*)ctx->slot[S_CURSORS],
                                h_stats, chunks, 0, max_len, d_bin_off)) return 1;
            if (h_stats[3] != 0) return fail (ctx, "partition kernel dropped %llu records in the exact layout", (unsigned long long)h_stats[3]);
        }
    }
    const void* src_bins[1] = { ctx->slot[S_COARSE] };
    const uint32_t* src_cur[1] = { (const uint32_t*)ctx->slot[S_CURSORS] };
    const uint64_t* src_off[1] = { d_bin_off };
    if (count_bins_impl (ctx, p, &g, 1, src_bins, src_cur, g.nb1, repart_host, total_kmers, out, to_host, d_bin_off ? src_off : 0)) return 1;
    out->stats[GATB_STAT_KMERS_VALID] = h_stats[0]; out->stats[GATB_STAT_KMERS_INVALID] = h_stats[1];
    out->stats[GATB_STAT_SEQUENCES] = n_reads; out->stats[GATB_STAT_NUCLEOTIDES] = total_nt; out->stats[GATB_STAT_RETRIES] = retries;
    float ms;
    cudaEventElapsedTime (&ms, ctx->ev[1], ctx->ev[2]); out->seconds[1] = ms * 1e-3;
    cudaEventElapsedTime (&ms, ctx->ev[1], ctx->ev[5]); out->seconds[6] = ms * 1e-3;
    cudaEventElapsedTime (&ms, ctx->kev[0], ctx->kev[1]); out->kernel_seconds[0] = ms * 1e-3;
    return 0;
}

extern "C" {

int gatb_gpu_count_dev (gatb_gpu_ctx* ctx, const gatb_gpu_params* p, const uint16_t* repart_table, const uint32_t* freq_order,
                        const uint8_t* d_packed_reads, const uint64_t* d_read_offsets_nt, uint64_t n_reads,
                        const uint32_t* d_n_mask, gatb_gpu_result* out)
{
    if (!ctx) return 1;
    cudaSetDevice (ctx->device);
    if (!out) return fail (ctx, "out is NULL");
    if (check_params (ctx, p, repart_table)) return 1;
    (void)freq_order;
    return count_dev_impl (ctx, p, repart_table, d_packed_reads, d_read_offsets_nt, n_reads, d_n_mask, out);
}

// ---- staged entry points for multi-GPU runs (one process per GPU; the exchange between the stages is the caller's) ----
int gatb_gpu_plan (gatb_gpu_ctx* ctx, const gatb_gpu_params* p, uint64_t total_kmers, uint64_t n_reads, int n_ranks, gatb_gpu_geometry* g)
{
    if (!ctx) return 1;
    if (check_params (ctx, p, (const uint16_t*)1)) return 1;
    return plan_geometry (ctx, p, total_kmers, n_reads, n_ranks, g);
}
int gatb_gpu_partition_range_into (gatb_gpu_ctx* ctx, const gatb_gpu_params* p, const gatb_gpu_geometry* g,
                                   const uint8_t* d_packed_reads, const uint64_t* d_read_offsets_nt, uint64_t first_read, uint64_t n_reads,
                                   const uint32_t* d_n_mask, void* d_bins, uint32_t* d_cursors, uint64_t* stats4)
{
    if (!ctx) return 1;
    cudaSetDevice (ctx->device);
    if (check_params (ctx, p, (const uint16_t*)1)) return 1;
    if (check_geometry (ctx, p, g)) return 1;
    uint64_t range_max_len = 0;
    {   // the limits of the partition kernels are enforced here too (the event word keeps 21, oriented 20, bits of k-mer index)
        uint64_t tk = 0, tn = 0, max_len = 0;
        if (workload_size (ctx, p, d_read_offsets_nt ? d_read_offsets_nt + first_read : 0, n_reads, &tk, &tn, &max_len)) return 1;
        if (max_len >= (1ULL << 21)) return fail (ctx, "reads longer than 2^21-1 nucleotides are not supported by the partition kernel yet (longest: %llu)", (unsigned long long)max_len);
        if (max_len >= (1ULL << 20) && !(p->path_flags & GATB_PATH_K1_GENERAL))
            return fail (ctx, "reads of 2^20 nucleotides and more need path_flags |= GATB_PATH_K1_GENERAL on every rank (plan, partition and count)");
        range_max_len = max_len;
    }
    unsigned long long h[4];
    if (partition_impl (ctx, p, g, d_packed_reads, d_read_offsets_nt, n_reads, d_n_mask, d_bins, d_cursors, h, 0, first_read, range_max_len)) return 1;
    for (int i = 0; i < 4; i++) stats4[i] = h[i];
    return 0;
}
int gatb_gpu_partition_into (gatb_gpu_ctx* ctx, const gatb_gpu_params* p, const gatb_gpu_geometry* g,
                             const uint8_t* d_packed_reads, const uint64_t* d_read_offsets_nt, uint64_t n_reads, const uint32_t* d_n_mask,
                             void* d_bins, uint32_t* d_cursors, uint64_t* stats4)
{ return gatb_gpu_partition_range_into (ctx, p, g, d_packed_reads, d_read_offsets_nt, 0, n_reads, d_n_mask, d_bins, d_cursors, stats4); }
int gatb_gpu_count_bins (gatb_gpu_ctx* ctx, const gatb_gpu_params* p, const gatb_gpu_geometry* g, int n_src,
                         const void* const* d_src_bins, const uint32_t* const* d_src_cursors,
                         uint32_t nb1_local, const uint16_t* repart_table, uint64_t kmers_bound, gatb_gpu_result* out)
{
    if (!ctx) return 1;
    cudaSetDevice (ctx->device);
    if (!out) return fail (ctx, "out is NULL");
    if (check_params (ctx, p, repart_table)) return 1;
    if (check_geometry (ctx, p, g)) return 1;
    if (nb1_local > g->nb1) return fail (ctx, "nb1_local %u exceeds the geometry's %u coarse bins", nb1_local, g->nb1);
    cudaEventRecord (ctx->ev[1], ctx->stream);
    if (count_bins_impl (ctx, p, g, n_src, d_src_bins, d_src_cursors, nb1_local, repart_table, kmers_bound, out)) return 1;
    float ms; cudaEventElapsedTime (&ms, ctx->ev[1], ctx->ev[5]); out->seconds[6] = ms * 1e-3;
    return 0;
}

int gatb_gpu_count_bins_routed (gatb_gpu_ctx* ctx, const gatb_gpu_params* p, const gatb_gpu_geometry* g, int n_src,
                                const void* const* d_src_bins, const uint32_t* const* d_src_cursors,
                                uint32_t nb1_local, const uint16_t* repart_table, uint64_t kmers_bound, int n_ranks,
                                uint64_t* send_counts, uint16_t** d_keys, gatb_gpu_result* out)
{
    if (!ctx) return 1;
    cudaSetDevice (ctx->device);
    if (!out || !send_counts || !d_keys) return fail (ctx, "out, send_counts or d_keys is NULL");
    if (check_params (ctx, p, repart_table)) return 1;
    if (check_geometry (ctx, p, g)) return 1;
    if (nb1_local > g->nb1) return fail (ctx, "nb1_local %u exceeds the geometry's %u coarse bins", nb1_local, g->nb1);
    if (n_ranks < 1 || n_ranks > GATB_GPU_MAX_RANKS) return fail (ctx, "n_ranks must be in [1,%d]", GATB_GPU_MAX_RANKS);
    cudaEventRecord (ctx->ev[1], ctx->stream);
    if (count_bins_impl (ctx, p, g, n_src, d_src_bins, d_src_cursors, nb1_local, repart_table, kmers_bound, out, false, 0, n_ranks, send_counts, d_keys)) return 1;
    float ms; cudaEventElapsedTime (&ms, ctx->ev[1], ctx->ev[5]); out->seconds[6] = ms * 1e-3;
    return 0;
}

int gatb_gpu_sort_routed (gatb_gpu_ctx* ctx, const gatb_gpu_params* p, const uint64_t* d_lo, const uint64_t* d_hi, const uint32_t* d_counts,
                          const uint16_t* d_keys, uint64_t n_items, int n_ranks, int rank, gatb_gpu_result* out)
{
    if (!ctx) return 1;
    cudaSetDevice (ctx->device);
    if (!out) return fail (ctx, "out is NULL");
    if (p->kmer_size < 1 || p->kmer_size > 63 || p->nb_partitions < 1 || p->nb_passes < 1 || p->histo_max < 1) return fail (ctx, "bad parameters");
    if (n_items && (!d_lo || !d_counts || !d_keys || (p->kmer_size >= 32 && !d_hi))) return fail (ctx, "item arrays are NULL");
    cudaEventRecord (ctx->ev[4], ctx->stream);
    if (n_ranks < 1 || n_ranks > GATB_GPU_MAX_RANKS || rank < 0 || rank >= n_ranks) return fail (ctx, "bad rank %d of %d", rank, n_ranks);
    // the keys this rank owns are rank, rank + n_ranks, ...: the items carry their index in that list (key / n_ranks)
    const uint64_t n_keys_all = (uint64_t)p->nb_partitions * p->nb_passes;
    const uint64_t n_keys_local = n_keys_all > (uint64_t)rank ? (n_keys_all - rank + n_ranks - 1) / n_ranks : 0;
    SortOut so;
    if (sort_stage (ctx, p, d_lo, d_hi, d_counts, d_keys, n_items, n_items, 0, false, &so, n_keys_local ? n_keys_local : 1)) return 1;
    {   // offsets of ALL keys: a key this rank does not own is empty (it starts where the next owned key starts)
        std::vector<uint64_t> loc (n_keys_local + 2, 0), all (n_keys_all + 1);
        CK (cudaMemcpyAsync (loc.data (), so.dr.offs, ((n_keys_local ? n_keys_local : 1) + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK (cudaStreamSynchronize (ctx->stream));
        for (uint64_t key = 0; key <= n_keys_all; key++)
        {
            uint64_t j = key > (uint64_t)rank ? (key - rank + n_ranks - 1) / n_ranks : 0;          // owned keys below 'key'
            if (j > n_keys_local) j = n_keys_local;
            all[key] = n_keys_local ? loc[j] : 0;
        }
        CK (cudaMemcpyAsync (so.dr.offs, all.data (), (n_keys_all + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK (cudaStreamSynchronize (ctx->stream));
    }
    cudaEventRecord (ctx->ev[5], ctx->stream);
    CK (cudaStreamSynchronize (ctx->stream));
    memset (out, 0, sizeof(*out));
    out->n_keys = (uint64_t)p->nb_partitions * p->nb_passes; out->n_items = n_items; out->on_device = 1;
    out->part_offsets = (uint64_t*)so.dr.offs; out->kmers_lo = (uint64_t*)so.dr.lo; out->kmers_hi = (uint64_t*)so.dr.hi;
    out->counts = (int32_t*)so.dr.cnt; out->histogram = (uint64_t*)so.dr.histo;
    float ms;
    cudaEventElapsedTime (&ms, ctx->ev[4], ctx->ev[5]); out->seconds[4] = ms * 1e-3;
    cudaEventElapsedTime (&ms, ctx->kev[6], ctx->kev[7]); out->kernel_seconds[3] = ms * 1e-3;
    out->kernel_seconds[5] = (double)so.n_big; out->kernel_seconds[6] = so.two_pass; out->kernel_seconds[7] = so.t_bits;      // diagnostics of the sort stage
    return 0;
}

void gatb_gpu_result_free (gatb_gpu_ctx* ctx, gatb_gpu_result* r)
{
    if (!r) return;
    if (ctx) cudaSetDevice (ctx->device);
    /* device arrays live in context slots, host arrays in the context's pinned staging buffer: nothing to free;
       both stay valid until the next count on the same context */
    memset (r, 0, sizeof(*r));
}

int gatb_gpu_count (gatb_gpu_ctx* ctx, const gatb_gpu_params* p, const uint16_t* repart_table, const uint32_t* freq_order,
                    const uint8_t* packed_reads, const uint64_t* read_offsets_nt, uint64_t n_reads, const uint32_t* n_mask,
                    gatb_gpu_result* out)
{
    if (!ctx) return 1;
    cudaSetDevice (ctx->device);
    if (!out) return fail (ctx, "out is NULL");
    if (check_params (ctx, p, repart_table)) return 1;
    if (!read_offsets_nt && p->read_len <= 0) return fail (ctx, "read_offsets_nt is NULL and read_len <= 0");
    (void)freq_order;
    cudaEventRecord (ctx->ev[0], ctx->stream);
    const uint64_t total_nt = read_offsets_nt ? read_offsets_nt[n_reads] : n_reads * (uint64_t)p->read_len;
    const uint64_t bytes = (total_nt + 3) / 4;
    if (ensure (ctx, S_READS, bytes + 64)) return 1;
    uint8_t* d_reads = (uint8_t*)ctx->slot[S_READS];
    const uint64_t* d_off = 0; const uint32_t* d_mask = 0;
    if (read_offsets_nt)
    {
        if (ensure (ctx, S_OFFSETS, (n_reads + 1) * 8)) return 1;
        CK (cudaMemcpyAsync (ctx->slot[S_OFFSETS], read_offsets_nt, (n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        d_off = (const uint64_t*)ctx->slot[S_OFFSETS];
    }
    if (n_mask)
    {
        uint64_t mw = (total_nt + 31) / 32;
        if (ensure (ctx, S_NMASK, mw * 4 + 16)) return 1;
        CK (cudaMemsetAsync ((uint8_t*)ctx->slot[S_NMASK] + mw * 4, 0, 16, ctx->stream));
        CK (cudaMemcpyAsync (ctx->slot[S_NMASK], n_mask, mw * 4, cudaMemcpyHostToDevice, ctx->stream));
        d_mask = (const uint32_t*)ctx->slot[S_NMASK];
    }
    // the packed reads travel in chunks on the copy stream; k1 of chunk c starts as soon as chunk c has landed
    ReadChunks rc; rc.n = (bytes > (64u << 20)) ? 8 : 1; rc.ready = ctx->cev;
    CK (cudaMemsetAsync (d_reads + (bytes & ~15ULL), 0, (bytes & 15) + 48, ctx->stream));      // zero the padding
    CK (cudaEventRecord (ctx->cev[39], ctx->stream));
    CK (cudaStreamWaitEvent (ctx->copy_stream, ctx->cev[39], 0));
    uint64_t byte_done = 0;
    for (int c = 0; c < rc.n; c++)
    {
        rc.first[c] = n_reads * (uint64_t)c / rc.n;
        const uint64_t last = n_reads * (uint64_t)(c + 1) / rc.n;
        const uint64_t end_nt = read_offsets_nt ? read_offsets_nt[last] : last * (uint64_t)p->read_len;
        uint64_t byte_end = (c == rc.n - 1) ? bytes : ((end_nt + 3) / 4 + 15) & ~15ULL;      // whole 16-byte units past the chunk's last read
        if (byte_end > bytes) byte_end = bytes;
        if (byte_end > byte_done) CK (cudaMemcpyAsync (d_reads + byte_done, packed_reads + byte_done, byte_end - byte_done, cudaMemcpyHostToDevice, ctx->copy_stream));
        byte_done = byte_end > byte_done ? byte_end : byte_done;
        CK (cudaEventRecord (ctx->cev[c], ctx->copy_stream));
    }
    rc.first[rc.n] = n_reads;
    gatb_gpu_result h;
    if (count_dev_impl (ctx, p, repart_table, d_reads, d_off, n_reads, d_mask, &h, &rc, true)) return 1;
    cudaEventRecord (ctx->ev[7], ctx->stream);
    CK (cudaStreamSynchronize (ctx->stream));
    float ms;
    cudaEventElapsedTime (&ms, ctx->ev[0], ctx->ev[7]); h.seconds[7] = ms * 1e-3;
    *out = h;
    return 0;
}

// =====================================================================================================================
//  several devices driven by ONE process (SURVEY.md 8b "gatb_gpu_create (n_gpus, dev_ids)"): the staged sequence -- plan, partition,
//  exchange of the coarse-bin regions, count -- with one host thread per device and peer copies instead of a collective library, then
//  the per-device ascending runs of every partition merged on the host: a C or C++ caller gets one result, as from gatb_gpu_count.
// =====================================================================================================================
struct MultiDev { gatb_gpu_ctx* ctx; uint64_t r0, r1; unsigned long long st[4]; int rc; gatb_gpu_result res; };

static int multi_partition (MultiDev* D, const gatb_gpu_params* p, const gatb_gpu_geometry* g, const uint8_t* packed, uint64_t bytes,
                            const uint64_t* offs, uint64_t n_reads, const uint32_t* n_mask, uint64_t total_nt)
{
    gatb_gpu_ctx* ctx = D->ctx;
    cudaSetDevice (ctx->device);
    if (ensure (ctx, S_READS, bytes + 64)) return 1;
    uint8_t* d_reads = (uint8_t*)ctx->slot[S_READS];
    const uint64_t nt_lo = offs ? offs[D->r0] : D->r0 * (uint64_t)p->read_len, nt_hi = offs ? offs[D->r1] : D->r1 * (uint64_t)p->read_len;
    // the device keeps the stream at its absolute offsets but receives only the bytes of its own reads (16-byte units; the
    // partition kernel looks up to 32 bytes past the last nucleotide)
    const uint64_t b_lo = (nt_lo / 4) & ~15ULL;
    uint64_t b_hi = (((nt_hi + 3) / 4) + 48 + 15) & ~15ULL; if (b_hi > bytes) b_hi = bytes;
    CK (cudaMemsetAsync (d_reads + (bytes & ~15ULL), 0, (bytes & 15) + 48, ctx->stream));
    if (b_hi > b_lo) CK (cudaMemcpyAsync (d_reads + b_lo, packed + b_lo, b_hi - b_lo, cudaMemcpyHostToDevice, ctx->stream));
    const uint64_t* d_off = 0; const uint32_t* d_mask = 0;
    if (offs)
    {
        if (ensure (ctx, S_OFFSETS, (n_reads + 1) * 8)) return 1;
        CK (cudaMemcpyAsync ((uint64_t*)ctx->slot[S_OFFSETS] + D->r0, offs + D->r0, (D->r1 - D->r0 + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        d_off = (const uint64_t*)ctx->slot[S_OFFSETS];
    }
    if (n_mask)
    {
        const uint64_t mw = (total_nt + 31) / 32;
        if (ensure (ctx, S_NMASK, mw * 4 + 16)) return 1;
        const uint64_t w_lo = nt_lo / 32; uint64_t w_hi = (nt_hi + 31) / 32 + 2; if (w_hi > mw) w_hi = mw;
        CK (cudaMemsetAsync ((uint8_t*)ctx->slot[S_NMASK] + mw * 4, 0, 16, ctx->stream));
        if (w_hi > w_lo) CK (cudaMemcpyAsync ((uint32_t*)ctx->slot[S_NMASK] + w_lo, n_mask + w_lo, (w_hi - w_lo) * 4, cudaMemcpyHostToDevice, ctx->stream));
        d_mask = (const uint32_t*)ctx->slot[S_NMASK];
    }
    if (ensure (ctx, S_COARSE, (size_t)g->nb1 * g->cap * g->record_bytes)) return 1;
    if (ensure (ctx, S_CURSORS, (size_t)g->nb1 * 4)) return 1;
    uint64_t st4[4];
    if (gatb_gpu_partition_range_into (ctx, p, g, d_reads, d_off, D->r0, D->r1 - D->r0, d_mask, ctx->slot[S_COARSE], (uint32_t*)ctx->slot[S_CURSORS], st4)) return 1;
    for (int i = 0; i < 4; i++) D->st[i] = st4[i];
    return 0;
}

int gatb_gpu_count_multi (gatb_gpu_ctx* const* ctxs, int n_dev, const gatb_gpu_params* p, const uint16_t* repart_table,
                          const uint8_t* packed_reads, const uint64_t* read_offsets_nt, uint64_t n_reads, const uint32_t* n_mask,
                          gatb_gpu_result* out)
{
    if (!ctxs || n_dev < 1 || !ctxs[0]) return 1;
    gatb_gpu_ctx* ctx = ctxs[0];                                     // errors are reported on the first context
    if (n_dev > GATB_GPU_MAX_RANKS) return fail (ctx, "at most %d devices", GATB_GPU_MAX_RANKS);
    for (int d = 0; d < n_dev; d++) if (!ctxs[d]) return fail (ctx, "context %d is NULL", d);
    if (!out) return fail (ctx, "out is NULL");
    if (n_dev == 1) return gatb_gpu_count (ctx, p, repart_table, 0, packed_reads, read_offsets_nt, n_reads, n_mask, out);
    cudaSetDevice (ctx->device);
    if (check_params (ctx, p, repart_table)) return 1;
    if (!read_offsets_nt && p->read_len <= 0) return fail (ctx, "read_offsets_nt is NULL and read_len <= 0");
    const int k = p->kmer_size, W = (k < 32) ? 1 : 2;
    const uint64_t n_keys = (uint64_t)p->nb_partitions * p->nb_passes;
    const uint64_t total_nt = read_offsets_nt ? read_offsets_nt[n_reads] : n_reads * (uint64_t)p->read_len;
    const uint64_t bytes = (total_nt + 3) / 4;
    uint64_t total_kmers = 0;
    if (read_offsets_nt) { for (uint64_t i = 0; i < n_reads; i++) { const uint64_t len = read_offsets_nt[i+1] - read_offsets_nt[i]; if (len >= (uint64_t)k) total_kmers += len - k + 1; } }
    else if (p->read_len >= k) total_kmers = n_reads * (uint64_t)(p->read_len - k + 1);
    gatb_gpu_geometry g;
    if (plan_geometry (ctx, p, total_kmers, n_reads, n_dev, &g)) return 1;
    std::vector<MultiDev> D (n_dev);
    for (int d = 0; d < n_dev; d++) { D[d].ctx = ctxs[d]; D[d].r0 = (n_reads * (uint64_t)d / n_dev) & ~31ULL; D[d].rc = 0; memset (&D[d].res, 0, sizeof(D[d].res)); }
    for (int d = 0; d < n_dev; d++) D[d].r1 = (d + 1 < n_dev) ? D[d + 1].r0 : n_reads;
    for (int d = 0; d < n_dev; d++) for (int e = 0; e < n_dev; e++) if (e != d)
    {   // direct peer copies where the devices allow them (cudaMemcpyPeer works either way)
        cudaSetDevice (ctxs[d]->device);
        int can = 0; cudaDeviceCanAccessPeer (&can, ctxs[d]->device, ctxs[e]->device);
        if (can) { cudaError_t pe = cudaDeviceEnablePeerAccess (ctxs[e]->device, 0); if (pe != cudaSuccess) cudaGetLastError (); }
    }
    // ---- stage 1: every device partitions its slice of the reads into all regions (one host thread per device) ----
    for (int attempt = 0; ; attempt++)
    {
        std::vector<std::thread> th;
        for (int d = 0; d < n_dev; d++) th.emplace_back ([&, d] () { D[d].rc = multi_partition (&D[d], p, &g, packed_reads, bytes, read_offsets_nt, n_reads, n_mask, total_nt); });
        for (auto& t : th) t.join ();
        for (int d = 0; d < n_dev; d++) if (D[d].rc) return fail (ctx, "device %d: %s", ctxs[d]->device, ctxs[d]->error.c_str ());
        unsigned long long dropped = 0; for (int d = 0; d < n_dev; d++) dropped += D[d].st[3];
        if (!dropped) break;
        if (attempt >= 1) return fail (ctx, "a coarse bin overflowed twice");
        // a bin overflowed somewhere: every device runs again with room for the largest demand seen
        uint32_t need = g.cap;
        for (int d = 0; d < n_dev; d++)
        {
            cudaSetDevice (ctxs[d]->device);
            std::vector<uint32_t> cur (g.nb1);
            if (cudaMemcpy (cur.data (), ctxs[d]->slot[S_CURSORS], (size_t)g.nb1 * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return fail (ctx, "cursor copy failed");
            for (uint32_t b = 0; b < g.nb1; b++) if (cur[b] > need) need = cur[b];
        }
        g.cap = (need + COARSE_BLK - 1) / COARSE_BLK * COARSE_BLK;
    }
    // ---- stage 2: region d of every device goes to device d (peer copies; the whole region: its unused rounds are slack) ----
    const size_t region_bytes = (size_t)g.bins_per_rank * g.cap * g.record_bytes, cur_bytes = (size_t)g.bins_per_rank * 4;
    for (int d = 0; d < n_dev; d++)
    {
        cudaSetDevice (ctxs[d]->device);
        if (ensure (ctxs[d], S_MISC, region_bytes * (n_dev - 1))) return fail (ctx, "device %d: %s", ctxs[d]->device, ctxs[d]->error.c_str ());
        if (ensure (ctxs[d], S_MISC2, cur_bytes * (n_dev - 1))) return fail (ctx, "device %d: %s", ctxs[d]->device, ctxs[d]->error.c_str ());
    }
    for (int d = 0; d < n_dev; d++)
    {
        cudaSetDevice (ctxs[d]->device);
        int slot = 0;
        for (int s = 0; s < n_dev; s++) if (s != d)
        {
            CK (cudaMemcpyPeerAsync ((uint8_t*)ctxs[d]->slot[S_MISC] + region_bytes * slot, ctxs[d]->device,
                                     (const uint8_t*)ctxs[s]->slot[S_COARSE] + region_bytes * d, ctxs[s]->device, region_bytes, ctxs[d]->stream));
            CK (cudaMemcpyPeerAsync ((uint8_t*)ctxs[d]->slot[S_MISC2] + cur_bytes * slot, ctxs[d]->device,
                                     (const uint8_t*)ctxs[s]->slot[S_CURSORS] + cur_bytes * d, ctxs[s]->device, cur_bytes, ctxs[d]->stream));
            slot++;
        }
    }
    for (int d = 0; d < n_dev; d++) { cudaSetDevice (ctxs[d]->device); CK (cudaStreamSynchronize (ctxs[d]->stream)); }
    // ---- stage 3: every device counts the bins it owns from all sources; results into its pinned host buffer ----
    {
        std::vector<std::thread> th;
        for (int d = 0; d < n_dev; d++) th.emplace_back ([&, d] ()
        {
            gatb_gpu_ctx* c = ctxs[d];
            cudaSetDevice (c->device);
            const void* bins[GATB_GPU_MAX_RANKS]; const uint32_t* curs[GATB_GPU_MAX_RANKS];
            int slot = 0;
            for (int s = 0; s < n_dev; s++)
            {
                if (s == d) { bins[s] = (const uint8_t*)c->slot[S_COARSE] + region_bytes * d; curs[s] = (const uint32_t*)((const uint8_t*)c->slot[S_CURSORS] + cur_bytes * d); }
                else        { bins[s] = (const uint8_t*)c->slot[S_MISC] + region_bytes * slot; curs[s] = (const uint32_t*)((const uint8_t*)c->slot[S_MISC2] + cur_bytes * slot); slot++; }
            }
            cudaEventRecord (c->ev[1], c->stream);
            D[d].rc = count_bins_impl (c, p, &g, n_dev, bins, curs, g.bins_per_rank, repart_table, total_kmers, &D[d].res, true);
        });
        for (auto& t : th) t.join ();
        for (int d = 0; d < n_dev; d++) if (D[d].rc) return fail (ctx, "device %d: %s", ctxs[d]->device, ctxs[d]->error.c_str ());
    }
    // ---- stage 4 (host): a k-mer lives on one device, so the runs of a partition are disjoint: merged into one ascending sequence ----
    uint64_t n_items = 0; for (int d = 0; d < n_dev; d++) n_items += D[d].res.n_items;
    const size_t hist_bytes = (size_t)(p->histo_max + 1) * 8, off_bytes = (n_keys + 1) * 8;
    const size_t need = off_bytes + hist_bytes + (n_items + 1) * (8 * W + 4) + 64;
    if (ctx->multi_cap < need) { free (ctx->multi_buf); ctx->multi_buf = malloc (need); ctx->multi_cap = ctx->multi_buf ? need : 0; }
    if (!ctx->multi_buf) return fail (ctx, "host allocation of the merged result (%zu bytes) failed", need);
    uint8_t* mb = (uint8_t*)ctx->multi_buf;
    uint64_t* m_lo = (uint64_t*)mb; mb += (n_items + 1) * 8;
    uint64_t* m_hi = 0; if (W == 2) { m_hi = (uint64_t*)mb; mb += (n_items + 1) * 8; }
    uint64_t* m_off = (uint64_t*)mb; mb += off_bytes;
    uint64_t* m_hist = (uint64_t*)mb; mb += hist_bytes;
    int32_t* m_cnt = (int32_t*)mb;
    uint64_t at = 0;
    for (uint64_t key = 0; key < n_keys; key++)
    {
        m_off[key] = at;
        uint64_t pos[GATB_GPU_MAX_RANKS], end[GATB_GPU_MAX_RANKS];
        for (int d = 0; d < n_dev; d++) { pos[d] = D[d].res.part_offsets[key]; end[d] = D[d].res.part_offsets[key + 1]; }
        for (;;)
        {
            int best = -1;
            for (int d = 0; d < n_dev; d++)
            {
                if (pos[d] >= end[d]) continue;
                if (best < 0) { best = d; continue; }
                const uint64_t bl = D[best].res.kmers_lo[pos[best]], dl = D[d].res.kmers_lo[pos[d]];
                if (W == 2)
                {
                    const uint64_t bh = D[best].res.kmers_hi[pos[best]], dh = D[d].res.kmers_hi[pos[d]];
                    if (dh < bh || (dh == bh && dl < bl)) best = d;
                }
                else if (dl < bl) best = d;
            }
            if (best < 0) break;
            m_lo[at] = D[best].res.kmers_lo[pos[best]]; if (W == 2) m_hi[at] = D[best].res.kmers_hi[pos[best]];
            m_cnt[at] = D[best].res.counts[pos[best]]; pos[best]++; at++;
        }
    }
    m_off[n_keys] = at;
    memset (m_hist, 0, hist_bytes);
    for (int d = 0; d < n_dev; d++) for (int i = 0; i <= p->histo_max; i++) m_hist[i] += D[d].res.histogram[i];
    memset (out, 0, sizeof(*out));
    out->n_keys = n_keys; out->n_items = n_items; out->part_offsets = m_off; out->kmers_lo = m_lo; out->kmers_hi = m_hi; out->counts = m_cnt; out->histogram = m_hist;
    out->on_device = 0;
    for (int d = 0; d < n_dev; d++)
    {
        out->stats[GATB_STAT_KMERS_VALID] += D[d].st[0]; out->stats[GATB_STAT_KMERS_INVALID] += D[d].st[1];
        for (int i : { (int)GATB_STAT_DISTINCT, (int)GATB_STAT_SOLID, (int)GATB_STAT_RECORDS, (int)GATB_STAT_BINS, (int)GATB_STAT_OVERFLOW_BINS, (int)GATB_STAT_RECORD_BYTES, (int)GATB_STAT_UNIQUE_RECORDS, 11, 12 })
            out->stats[i] += D[d].res.stats[i];
        for (int i = 0; i < 8; i++) { if (D[d].res.seconds[i] > out->seconds[i]) out->seconds[i] = D[d].res.seconds[i]; if (i < 5 && D[d].res.kernel_seconds[i] > out->kernel_seconds[i]) out->kernel_seconds[i] = D[d].res.kernel_seconds[i]; }
    }
    out->stats[GATB_STAT_SEQUENCES] = n_reads; out->stats[GATB_STAT_NUCLEOTIDES] = total_nt;
    return 0;
}

// =====================================================================================================================
//  streaming input
// =====================================================================================================================
// grows a slot to 'bytes' keeping its first 'used' bytes; everything beyond 'used' is zero afterwards
static int grow_keep (gatb_gpu_ctx* ctx, int s, size_t bytes, size_t used)
{
    if (ctx->slot_cap[s] >= bytes) return 0;
    size_t want = ctx->slot_cap[s] * 2; if (want < bytes) want = bytes; want = (want + 255) & ~(size_t)255;
    void* fresh = 0;
    cudaError_t e = cudaMalloc (&fresh, want);
    if (e != cudaSuccess) { want = (bytes + 255) & ~(size_t)255; e = cudaMalloc (&fresh, want); }
    if (e != cudaSuccess) return fail (ctx, "cudaMalloc of %zu bytes (slot %d) failed: %s", want, s, cudaGetErrorString (e));
    if (used) CK (cudaMemcpyAsync (fresh, ctx->slot[s], used, cudaMemcpyDeviceToDevice, ctx->stream));
    CK (cudaMemsetAsync ((uint8_t*)fresh + used, 0, want - used, ctx->stream));
    CK (cudaStreamSynchronize (ctx->stream));
    if (ctx->slot[s]) cudaFree (ctx->slot[s]);
    ctx->slot[s] = fresh; ctx->slot_cap[s] = want;
    return 0;
}

int gatb_gpu_reads_begin (gatb_gpu_ctx* ctx, uint64_t expected_nt)
{
    if (!ctx) return 1;
    cudaSetDevice (ctx->device);
    ctx->push_nt = ctx->push_seqs = ctx->push_invalid = 0;
    if (grow_keep (ctx, S_READS, expected_nt / 4 + 256, 0)) return 1;
    if (grow_keep (ctx, S_NMASK, expected_nt / 8 + 256, 0)) return 1;
    if (grow_keep (ctx, S_OFFSETS, 4096, 0)) return 1;
    // appended batches OR their edge words in: the buffers start from zero
    CK (cudaMemsetAsync (ctx->slot[S_READS], 0, ctx->slot_cap[S_READS], ctx->stream));
    CK (cudaMemsetAsync (ctx->slot[S_NMASK], 0, ctx->slot_cap[S_NMASK], ctx->stream));
    CK (cudaMemsetAsync (ctx->slot[S_OFFSETS], 0, 8, ctx->stream));
    return 0;
}

int gatb_gpu_reads_push_ascii (gatb_gpu_ctx* ctx, const char* ascii, const uint64_t* seq_offsets, uint64_t n_seqs)
{
    if (!ctx) return 1;
    cudaSetDevice (ctx->device);
    if (n_seqs == 0) return 0;
    if (!ascii || !seq_offsets) return fail (ctx, "reads_push_ascii: NULL buffer");
    const uint64_t n = seq_offsets[n_seqs] - seq_offsets[0];
    for (uint64_t i = 0; i < n_seqs; i++) if (seq_offsets[i+1] < seq_offsets[i]) return fail (ctx, "reads_push_ascii: offsets must not decrease");
    const uint64_t base = ctx->push_nt, total = base + n;
    if (grow_keep (ctx, S_READS, total / 4 + 256, base / 4 + 8)) return 1;
    if (grow_keep (ctx, S_NMASK, total / 8 + 256, base / 8 + 8)) return 1;
    if (grow_keep (ctx, S_OFFSETS, (ctx->push_seqs + n_seqs + 1) * 8 + 64, (ctx->push_seqs + 1) * 8)) return 1;
    if (ensure (ctx, S_MISC, n + (n_seqs + 1) * 8 + 64)) return 1;
    if (ensure (ctx, S_STATS, 64 * 8)) return 1;
    uint64_t* d_in_off = (uint64_t*)ctx->slot[S_MISC];
    char* d_ascii = (char*)ctx->slot[S_MISC] + (n_seqs + 1) * 8;
    unsigned long long* d_bad = (unsigned long long*)ctx->slot[S_STATS] + 40;
    CK (cudaMemcpyAsync (d_in_off, seq_offsets, (n_seqs + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (n) CK (cudaMemcpyAsync (d_ascii, ascii + seq_offsets[0], n, cudaMemcpyHostToDevice, ctx->stream));
    CK (cudaMemsetAsync (d_bad, 0, 8, ctx->stream));
    CK (launch_pack_ascii_at (lctx (ctx), d_ascii, n, base, (uint32_t*)ctx->slot[S_READS], (uint32_t*)ctx->slot[S_NMASK], d_bad));
    CK (launch_rebase_offsets (lctx (ctx), d_in_off, n_seqs + 1, base, (uint64_t*)ctx->slot[S_OFFSETS] + ctx->push_seqs));
    unsigned long long bad = 0;
    CK (cudaMemcpyAsync (&bad, d_bad, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK (cudaStreamSynchronize (ctx->stream));                    // the caller may reuse its buffers on return
    ctx->push_invalid += bad; ctx->push_nt = total; ctx->push_seqs += n_seqs;
    return 0;
}

// FASTA / FASTQ text parsed on the device (k_parse.cu): lines, records, offsets, packing
int gatb_gpu_reads_push_text (gatb_gpu_ctx* ctx, const char* text, uint64_t n, int format)
{
    if (!ctx) return 1;
    cudaSetDevice (ctx->device);
    if (n == 0) return 0;
    if (!text) return fail (ctx, "reads_push_text: NULL buffer");
    if (format != GATB_TEXT_FASTA && format != GATB_TEXT_FASTQ) return fail (ctx, "reads_push_text: unknown format %d", format);
    if (n >= (1ULL << 32)) return fail (ctx, "reads_push_text: batches of 4 GiB and more are not supported (cut the file into smaller batches)");
    LaunchCtx L = lctx (ctx);
    const uint64_t nblk = text_blocks (n);
    // ---- text to the device, newline census ----
    if (ensure (ctx, S_MISC, n + 64)) return 1;
    if (ensure (ctx, S_TOTCUR, nblk * 4 + 64)) return 1;
    if (ensure (ctx, S_COARSEOFF, (nblk + 1) * 8 + 64)) return 1;
    if (ensure (ctx, S_SCAN, scan_scratch_elems (nblk > (n / 16 + 2) ? nblk : (n / 16 + 2)) * 8)) return 1;
    if (ensure (ctx, S_STATS, 64 * 8)) return 1;
    char* d_text = (char*)ctx->slot[S_MISC];
    CK (cudaMemcpyAsync (d_text, text, n, cudaMemcpyHostToDevice, ctx->stream));
    CK (launch_text_count_newlines (L, d_text, n, (uint32_t*)ctx->slot[S_TOTCUR]));
    CK (launch_scan_u32_to_u64 (L, (const uint32_t*)ctx->slot[S_TOTCUR], (uint64_t*)ctx->slot[S_COARSEOFF], nblk, (uint64_t*)ctx->slot[S_SCAN]));
    uint64_t n_newlines = 0; char last = 0;
    CK (cudaMemcpyAsync (&n_newlines, (const uint64_t*)ctx->slot[S_COARSEOFF] + nblk, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK (cudaStreamSynchronize (ctx->stream));
    last = text[n - 1];
    const uint64_t n_lines = n_newlines + (last == '\n' ? 0 : 1);
    if (n_lines == 0) return 0;
    // ---- line table: starts, sequence lengths, header flags; destinations and record indices by two scans ----
    if (ensure (ctx, S_BUCKETOFF, (n_lines + 2) * 8)) return 1;            // line_start
    if (ensure (ctx, S_BUCKETCNT, (n_lines + 2) * 4)) return 1;            // seq_len
    if (ensure (ctx, S_BUCKETOF,  (n_lines + 2) * 4)) return 1;            // is_header
    if (ensure (ctx, S_BIGLIST,   (n_lines + 2) * 8)) return 1;            // line_dst
    if (ensure (ctx, S_MISC2,     (n_lines + 2) * 8)) return 1;            // line_rec
    if (ensure (ctx, S_SCAN, scan_scratch_elems (n_lines + 1) * 8)) return 1;
    uint64_t* d_ls = (uint64_t*)ctx->slot[S_BUCKETOFF];
    CK (cudaMemsetAsync (d_ls, 0, 8, ctx->stream));
    CK (launch_text_line_starts (L, d_text, n, (const uint64_t*)ctx->slot[S_COARSEOFF], d_ls));
    uint32_t* d_len = (uint32_t*)ctx->slot[S_BUCKETCNT]; uint32_t* d_hdr = (uint32_t*)ctx->slot[S_BUCKETOF];
    uint64_t* d_dst = (uint64_t*)ctx->slot[S_BIGLIST];   uint64_t* d_rec = (uint64_t*)ctx->slot[S_MISC2];
    CK (launch_text_lines (L, d_text, n, d_ls, n_lines, format, d_len, d_hdr));
    CK (launch_scan_u32_to_u64 (L, d_len, d_dst, n_lines, (uint64_t*)ctx->slot[S_SCAN]));
    CK (launch_scan_u32_to_u64 (L, d_hdr, d_rec, n_lines, (uint64_t*)ctx->slot[S_SCAN]));
    uint64_t tot[2] = { 0, 0 };
    CK (cudaMemcpyAsync (&tot[0], d_dst + n_lines, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK (cudaMemcpyAsync (&tot[1], d_rec + n_lines, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK (cudaStreamSynchronize (ctx->stream));
    const uint64_t total_nt = tot[0], n_recs = tot[1];
    if (n_recs == 0) return total_nt ? fail (ctx, "reads_push_text: sequence data without a record header (batches must start at a record)") : 0;
    // ---- append: offsets, packed nucleotides, invalid mask ----
    const uint64_t base = ctx->push_nt, total = base + total_nt;
    if (grow_keep (ctx, S_READS, total / 4 + 256, base / 4 + 8)) return 1;
    if (grow_keep (ctx, S_NMASK, total / 8 + 256, base / 8 + 8)) return 1;
    if (grow_keep (ctx, S_OFFSETS, (ctx->push_seqs + n_recs + 1) * 8 + 64, (ctx->push_seqs + 1) * 8)) return 1;
    unsigned long long* d_bad = (unsigned long long*)ctx->slot[S_STATS] + 40;
    CK (cudaMemsetAsync (d_bad, 0, 8, ctx->stream));
    CK (launch_text_offsets (L, d_hdr, d_dst, d_rec, n_lines, base, (uint64_t*)ctx->slot[S_OFFSETS] + ctx->push_seqs, n_recs, total_nt));
    CK (launch_text_pack (L, d_text, d_ls, d_dst, n_lines, total_nt, base, (uint32_t*)ctx->slot[S_READS], (uint32_t*)ctx->slot[S_NMASK], d_bad));
    unsigned long long bad = 0;
    CK (cudaMemcpyAsync (&bad, d_bad, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK (cudaStreamSynchronize (ctx->stream));
    ctx->push_invalid += bad; ctx->push_nt = total; ctx->push_seqs += n_recs;
    return 0;
}

// what was pushed so far: [0] sequences [1] nucleotides [2] shortest [3] longest [4] sum of squared lengths (double) [5] invalid nucleotides
int gatb_gpu_reads_info (gatb_gpu_ctx* ctx, uint64_t* info6)
{
    if (!ctx) return 1;
    cudaSetDevice (ctx->device);
    if (!info6) return fail (ctx, "reads_info: NULL output");
    if (ensure (ctx, S_STATS, 64 * 8)) return 1;
    unsigned long long* d = (unsigned long long*)ctx->slot[S_STATS] + 44;
    unsigned long long h[3] = { ~0ULL, 0, 0 };
    CK (cudaMemcpyAsync (d, h, 24, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->push_seqs) CK (launch_text_stats (lctx (ctx), (const uint64_t*)ctx->slot[S_OFFSETS], ctx->push_seqs, d));
    CK (cudaMemcpyAsync (h, d, 24, cudaMemcpyDeviceToHost, ctx->stream));
    CK (cudaStreamSynchronize (ctx->stream));
    info6[0] = ctx->push_seqs; info6[1] = ctx->push_nt; info6[2] = ctx->push_seqs ? h[0] : 0; info6[3] = h[1]; info6[4] = h[2]; info6[5] = ctx->push_invalid;
    return 0;
}

int gatb_gpu_reads_count (gatb_gpu_ctx* ctx, const gatb_gpu_params* p, const uint16_t* repart_table, const uint32_t* freq_order, gatb_gpu_result* out)
{
    if (!ctx) return 1;
    cudaSetDevice (ctx->device);
    if (!out) return fail (ctx, "out is NULL");
    if (check_params (ctx, p, repart_table)) return 1;
    (void)freq_order;
    if (ctx->push_seqs == 0) { if (gatb_gpu_reads_begin (ctx, 0)) return 1; }
    gatb_gpu_params pp = *p; pp.read_len = 0;                    // pushed reads always come with offsets
    cudaEventRecord (ctx->ev[0], ctx->stream);
    gatb_gpu_result h;
    if (count_dev_impl (ctx, &pp, repart_table, (const uint8_t*)ctx->slot[S_READS], (const uint64_t*)ctx->slot[S_OFFSETS], ctx->push_seqs,
                        ctx->push_invalid ? (const uint32_t*)ctx->slot[S_NMASK] : 0, &h, 0, true)) return 1;
    cudaEventRecord (ctx->ev[7], ctx->stream);
    CK (cudaStreamSynchronize (ctx->stream));
    float ms; cudaEventElapsedTime (&ms, ctx->ev[0], ctx->ev[7]); h.seconds[7] = ms * 1e-3;
    *out = h;
    return 0;
}

// =====================================================================================================================
//  GATB-exact super-k-mers (rows A3-A6)
// =====================================================================================================================
// ---- Repartitor table (row f3): sampling pass on the device (k_repart.cu), distribution on the host (repart_host.h) ----
int gatb_gpu_repartition (gatb_gpu_ctx* ctx, const gatb_gpu_params* p, const uint8_t* packed_reads, const uint64_t* read_offsets_nt,
                          uint64_t n_reads, const uint32_t* n_mask, uint64_t nb_seqs_to_see, uint16_t* table_out, uint64_t* info3)
{
    if (!ctx) return 1;
    cudaSetDevice (ctx->device);
    if (!table_out) return fail (ctx, "table_out is NULL");
    const int k = p->kmer_size, m = p->minimizer_size;
    if (k < 2 || k > 63) return fail (ctx, "gatb_gpu_repartition: kmer_size %d not supported (1 < k <= 63)", k);
    if (m < 2 || m > 12 || m >= k) return fail (ctx, "Bad values for kmer %d and minimizer %d", k, m);
    if (p->nb_partitions < 1 || p->nb_partitions > 65535) return fail (ctx, "nb_partitions must be in [1,65535]");
    if (p->minimizer_type != 0) return fail (ctx, "minimizer_type %d (frequency order) is not supported on the device path yet", p->minimizer_type);
    if (!read_offsets_nt && p->read_len <= 0) return fail (ctx, "read_offsets_nt is NULL and read_len <= 0");
    LaunchCtx L = lctx (ctx);
    const uint64_t total_nt = read_offsets_nt ? read_offsets_nt[n_reads] : n_reads * (uint64_t)p->read_len;
    const uint64_t bytes = (total_nt + 3) / 4;
    if (ensure (ctx, S_READS, bytes + 64)) return 1;
    CK (cudaMemsetAsync ((uint8_t*)ctx->slot[S_READS] + (bytes & ~15ULL), 0, (bytes & 15) + 48, ctx->stream));
    CK (cudaMemcpyAsync (ctx->slot[S_READS], packed_reads, bytes, cudaMemcpyHostToDevice, ctx->stream));
    const uint64_t* d_off = 0; const uint32_t* d_mask = 0;
    if (read_offsets_nt)
    {
        if (ensure (ctx, S_OFFSETS, (n_reads + 1) * 8)) return 1;
        CK (cudaMemcpyAsync (ctx->slot[S_OFFSETS], read_offsets_nt, (n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        d_off = (const uint64_t*)ctx->slot[S_OFFSETS];
    }
    if (n_mask)
    {
        uint64_t mw = (total_nt + 31) / 32;
        if (ensure (ctx, S_NMASK, mw * 4 + 16)) return 1;
        CK (cudaMemsetAsync ((uint8_t*)ctx->slot[S_NMASK] + mw * 4, 0, 16, ctx->stream));
        CK (cudaMemcpyAsync (ctx->slot[S_NMASK], n_mask, mw * 4, cudaMemcpyHostToDevice, ctx->stream));
        d_mask = (const uint32_t*)ctx->slot[S_NMASK];
    }
    const uint64_t nb_minims = 1ULL << (2 * m);
    if (ensure (ctx, S_MISC, (n_reads + 1) * 4)) return 1;
    if (ensure (ctx, S_MISC2, nb_minims * 8)) return 1;
    // pass 1: super-k-mers per read; the reference's iteration is cancelled after the read with which their running number passes
    // nb_seqs_to_see (SampleRepart::processSuperkmer, RepartitionAlgorithm.cpp:204-209; the flag is looked at between two reads)
    CK (launch_repart_sample (L, (const uint64_t*)ctx->slot[S_READS], d_off, p->read_len, d_mask, n_reads, k, m, (uint32_t*)ctx->slot[S_MISC], 0));
    std::vector<uint32_t> per_read (n_reads);
    CK (cudaMemcpyAsync (per_read.data (), ctx->slot[S_MISC], n_reads * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK (cudaStreamSynchronize (ctx->stream));
    uint64_t n_sampled = n_reads, seen = 0;
    for (uint64_t r = 0; r < n_reads; r++) { seen += per_read[r]; if (seen > nb_seqs_to_see) { n_sampled = r + 1; break; } }
    // pass 2: kx-mers per minimizer over the sampled reads
    CK (cudaMemsetAsync (ctx->slot[S_MISC2], 0, nb_minims * 8, ctx->stream));
    CK (launch_repart_sample (L, (const uint64_t*)ctx->slot[S_READS], d_off, p->read_len, d_mask, n_sampled, k, m, 0, (unsigned long long*)ctx->slot[S_MISC2]));
    std::vector<unsigned long long> kx (nb_minims);
    CK (cudaMemcpyAsync (kx.data (), ctx->slot[S_MISC2], nb_minims * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK (cudaStreamSynchronize (ctx->stream));
    repartition_distribute (kx, p->nb_partitions, table_out);
    if (info3) { unsigned long long tot = 0; for (uint64_t i = 0; i < nb_minims; i++) tot += kx[i]; info3[0] = n_sampled; info3[1] = seen; info3[2] = tot; }
    return 0;
}

int gatb_gpu_superkmers (gatb_gpu_ctx* ctx, const gatb_gpu_params* p, const uint16_t* repart_table,
                         const uint8_t* packed_reads, const uint64_t* read_offsets_nt, uint64_t n_reads,
                         const uint32_t* n_mask, uint8_t** streams, uint64_t* stream_sizes, uint64_t* stats_out)
{
    if (!ctx) return 1;
    cudaSetDevice (ctx->device);
    if (check_params (ctx, p, repart_table)) return 1;
    if (!read_offsets_nt && p->read_len <= 0) return fail (ctx, "read_offsets_nt is NULL and read_len <= 0");
    LaunchCtx L = lctx (ctx);
    const int k = p->kmer_size, m = p->minimizer_size, W = (k < 32) ? 1 : 2;
    const uint64_t n_keys = (uint64_t)p->nb_partitions * p->nb_passes;
    const uint64_t total_nt = read_offsets_nt ? read_offsets_nt[n_reads] : n_reads * (uint64_t)p->read_len;
    const uint64_t bytes = (total_nt + 3) / 4;
    if (ensure (ctx, S_READS, bytes + 64)) return 1;
    CK (cudaMemsetAsync ((uint8_t*)ctx->slot[S_READS] + (bytes & ~15ULL), 0, (bytes & 15) + 48, ctx->stream));
    CK (cudaMemcpyAsync (ctx->slot[S_READS], packed_reads, bytes, cudaMemcpyHostToDevice, ctx->stream));
    const uint64_t* d_off = 0; const uint32_t* d_mask = 0;
    if (read_offsets_nt)
    {
        for (uint64_t i = 0; i < n_reads; i++) if (read_offsets_nt[i+1] - read_offsets_nt[i] >= (1ULL << 21)) return fail (ctx, "reads longer than 2^21-1 nucleotides are not supported yet");
        if (ensure (ctx, S_OFFSETS, (n_reads + 1) * 8)) return 1;
        CK (cudaMemcpyAsync (ctx->slot[S_OFFSETS], read_offsets_nt, (n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        d_off = (const uint64_t*)ctx->slot[S_OFFSETS];
    }
    if (n_mask)
    {
        uint64_t mw = (total_nt + 31) / 32;
        if (ensure (ctx, S_NMASK, mw * 4 + 16)) return 1;
        CK (cudaMemsetAsync ((uint8_t*)ctx->slot[S_NMASK] + mw * 4, 0, 16, ctx->stream));
        CK (cudaMemcpyAsync (ctx->slot[S_NMASK], n_mask, mw * 4, cudaMemcpyHostToDevice, ctx->stream));
        d_mask = (const uint32_t*)ctx->slot[S_NMASK];
    }
    if (ensure (ctx, S_REPART, (1ULL << (2*m)) * 2)) return 1;
    if (repart_table) CK (cudaMemcpyAsync (ctx->slot[S_REPART], repart_table, (1ULL << (2*m)) * 2, cudaMemcpyHostToDevice, ctx->stream));
    else              CK (cudaMemsetAsync (ctx->slot[S_REPART], 0, (1ULL << (2*m)) * 2, ctx->stream));
    if (ensure (ctx, S_CURSORS, n_keys * 4)) return 1;
    if (ensure (ctx, S_STATS, 64 * 8)) return 1;
    unsigned long long* d_stats = (unsigned long long*)ctx->slot[S_STATS];

    K1Params k1; memset (&k1, 0, sizeof(k1));
    k1.words = (const uint64_t*)ctx->slot[S_READS]; k1.offsets = d_off; k1.nmask = d_mask; k1.n_reads = n_reads; k1.read_len = p->read_len;
    k1.k = k; k1.m = m; k1.w = k - m + 1; k1.maxlen = (W == 1) ? 28 : 60;
    k1.mmask = (1u << (2*m)) - 1; k1.mask_ma1 = gatb_mask_ma1 (m);
    k1.mode = K1_MODE_GATB; k1.repart = (const uint16_t*)ctx->slot[S_REPART]; k1.nb_partitions = p->nb_partitions; k1.nb_passes = p->nb_passes;
    k1.nb1 = (uint32_t)n_keys; k1.fine_bits = 0; k1.cursors = (uint32_t*)ctx->slot[S_CURSORS]; k1.stats = d_stats;
    // pass 1: demand per key; pass 2: fill exactly
    k1.count_only = 1; k1.cap = 0; k1.bins = 0;
    CK (cudaMemsetAsync (k1.cursors, 0, n_keys * 4, ctx->stream));
    CK (cudaMemsetAsync (d_stats, 0, 4 * 8, ctx->stream));
    if (n_reads) CK (launch_k1 (L, k1));
    std::vector<uint32_t> cur (n_keys);
    CK (cudaMemcpyAsync (cur.data (), k1.cursors, n_keys * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK (cudaStreamSynchronize (ctx->stream));
    uint64_t cap = 8; for (uint64_t i = 0; i < n_keys; i++) if (cur[i] > cap) cap = cur[i];
    if (ensure (ctx, S_COARSE, n_keys * cap * 16 * W)) return 1;
    k1.count_only = 0; k1.cap = (uint32_t)cap; k1.bins = ctx->slot[S_COARSE];
    CK (cudaMemsetAsync (k1.cursors, 0, n_keys * 4, ctx->stream));
    CK (cudaMemsetAsync (d_stats, 0, 4 * 8, ctx->stream));
    if (n_reads) CK (launch_k1 (L, k1));
    unsigned long long h_stats[4];
    CK (cudaMemcpyAsync (h_stats, d_stats, 4 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    // serialise into the reference's byte format
    if (ensure (ctx, S_COUNTERS, (2 * n_keys + 2) * 8 + 16 * 8)) return 1;
    unsigned long long* d_bytes = (unsigned long long*)ctx->slot[S_COUNTERS]; unsigned long long* d_cur = d_bytes + n_keys;
    CK (cudaMemsetAsync (d_bytes, 0, 2 * n_keys * 8, ctx->stream));
    CK (launch_serialize_sizes (L, W, k, k1.bins, k1.cursors, (uint32_t)n_keys, (uint32_t)cap, d_bytes));
    std::vector<unsigned long long> kb (n_keys);
    CK (cudaMemcpyAsync (kb.data (), d_bytes, n_keys * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK (cudaStreamSynchronize (ctx->stream));
    std::vector<uint64_t> off (n_keys + 1, 0);
    for (uint64_t i = 0; i < n_keys; i++) off[i+1] = off[i] + kb[i];
    if (ensure (ctx, S_BUCKETOFF, (n_keys + 1) * 8)) return 1;
    if (ensure (ctx, S_FINE, off[n_keys] + 16)) return 1;
    CK (cudaMemcpyAsync (ctx->slot[S_BUCKETOFF], off.data (), (n_keys + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK (launch_serialize_write (L, W, k, k1.bins, k1.cursors, (uint32_t)n_keys, (uint32_t)cap, (const uint64_t*)ctx->slot[S_BUCKETOFF], d_cur, (uint8_t*)ctx->slot[S_FINE]));
    for (uint64_t i = 0; i < n_keys; i++)
    {
        streams[i] = (uint8_t*) malloc (kb[i] ? kb[i] : 1); stream_sizes[i] = kb[i];
        if (kb[i]) CK (cudaMemcpyAsync (streams[i], (uint8_t*)ctx->slot[S_FINE] + off[i], kb[i], cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK (cudaStreamSynchronize (ctx->stream));
    if (stats_out) { stats_out[0] = h_stats[2]; stats_out[1] = h_stats[0]; stats_out[2] = h_stats[0]; stats_out[3] = h_stats[1]; }
    return 0;
}
void gatb_gpu_free_host (void* p) { free (p); }

// =====================================================================================================================
//  Bloom
// =====================================================================================================================
static const double g_rvalues_col1[129] = GATB_RVALUES_COL1_INIT;

int gatb_gpu_bloom_params (int kmer_size, uint64_t nb_solid, uint64_t* bloom_size, int32_t* nb_hash)
{
    if (kmer_size < 0 || kmer_size > 128) return 1;
    float bits = (float) g_rvalues_col1[kmer_size];                    // DebloomAlgorithm.cpp:638
    if (bits == 0) bits = 1;                                            // :648
    uint64_t est = (uint64_t)(nb_solid * bits);                        // float32 product, BloomAlgorithm.cpp:162
    if (est == 0) est = 1000;                                           // :165
    *bloom_size = est; *nb_hash = (int32_t) floorf (0.7 * bits);        // :163
    return 0;
}
// BloomContainer ctor (Bloom.hpp:184-199) and BloomCacheCoherent ctor (:437-442)
static void bloom_layout (int kind, uint64_t bloom_size, uint64_t* nchar, uint64_t* tai_out, int* pow2_out, uint64_t* reduced_out)
{
    uint64_t tai = (kind == GATB_BLOOM_BASIC) ? bloom_size : bloom_size + 2 * 4096;
    *nchar = 1 + tai / 8;
    int pow2 = (tai && !(tai & (tai - 1)));
    if (pow2) tai--;
    *tai_out = tai; *pow2_out = pow2; *reduced_out = tai - 2 * 4096;
}
int gatb_gpu_bloom_layout (int kind, uint64_t bloom_size, uint64_t* nbytes, uint64_t* bit_size)
{
    if (kind < 0 || kind > 2) return 1;
    uint64_t nchar, tai, reduced; int pow2;
    bloom_layout (kind, bloom_size, &nchar, &tai, &pow2, &reduced);
    *nbytes = nchar; *bit_size = (kind == GATB_BLOOM_BASIC) ? tai : reduced;
    return 0;
}
int gatb_gpu_bloom_dev (gatb_gpu_ctx* ctx, int kind, uint64_t bloom_size, int nb_hash, int kmer_size,
                        const uint64_t* d_lo, const uint64_t* d_hi, uint64_t n, uint8_t* d_out_bytes)
{
    if (!ctx) return 1;
    cudaSetDevice (ctx->device);
    if (kind < 0 || kind > 2) return fail (ctx, "bad Bloom kind %d in createBloom", kind);          // Bloom.hpp:1263
    if (kmer_size < 3 || kmer_size > 63) return fail (ctx, "kmer_size %d not supported", kmer_size);
    if (nb_hash < 1 || nb_hash > 10) return fail (ctx, "nb_hash must be in [1,10]");
    const int W = kmer_size < 32 ? 1 : 2;
    if (W == 2 && !d_hi) return fail (ctx, "kmers_hi is required for kmer_size >= 32");
    uint64_t nchar, tai, reduced; int pow2;
    bloom_layout (kind, bloom_size, &nchar, &tai, &pow2, &reduced);
    if (kind != GATB_BLOOM_BASIC && bloom_size == 0) return fail (ctx, "bloom_size must be > 0");
    CK (cudaMemsetAsync (d_out_bytes, 0, (nchar + 3) & ~3ULL, ctx->stream));
    CK (launch_bloom_insert (lctx (ctx), kind, W, kmer_size, nb_hash, tai, pow2, reduced, d_lo, d_hi, n, (uint32_t*)d_out_bytes));
    return 0;
}
int gatb_gpu_bloom (gatb_gpu_ctx* ctx, int kind, uint64_t bloom_size, int nb_hash, int kmer_size,
                    const uint64_t* lo, const uint64_t* hi, uint64_t n, uint8_t* out_bytes)
{
    if (!ctx) return 1;
    cudaSetDevice (ctx->device);
    if (kind < 0 || kind > 2) return fail (ctx, "bad Bloom kind %d in createBloom", kind);
    const int W = kmer_size < 32 ? 1 : 2;
    if (W == 2 && !hi) return fail (ctx, "kmers_hi is required for kmer_size >= 32");
    uint64_t nchar, bits;
    gatb_gpu_bloom_layout (kind, bloom_size, &nchar, &bits);
    uint64_t na = n ? n : 1;
    if (ensure (ctx, S_MISC, na * 8 * W)) return 1;
    if (ensure (ctx, S_MISC2, nchar + 64)) return 1;
    uint64_t* d_lo = (uint64_t*)ctx->slot[S_MISC]; uint64_t* d_hi = (W == 2) ? d_lo + na : 0;
    if (n) { CK (cudaMemcpyAsync (d_lo, lo, n * 8, cudaMemcpyHostToDevice, ctx->stream)); if (W == 2) CK (cudaMemcpyAsync (d_hi, hi, n * 8, cudaMemcpyHostToDevice, ctx->stream)); }
    if (gatb_gpu_bloom_dev (ctx, kind, bloom_size, nb_hash, kmer_size, d_lo, d_hi, n, (uint8_t*)ctx->slot[S_MISC2])) return 1;
    CK (cudaMemcpyAsync (out_bytes, ctx->slot[S_MISC2], nchar, cudaMemcpyDeviceToHost, ctx->stream));
    CK (cudaStreamSynchronize (ctx->stream));
    return 0;
}

// =====================================================================================================================
//  Histogram cutoff: Histogram::compute_threshold, tools/misc/impl/Histogram.cpp:61-190.  A 10 001-entry table:
//  host arithmetic in doubles exactly as the reference (the device produced the table).
// =====================================================================================================================
int gatb_gpu_histogram_cutoff (const uint64_t* h, int histo_max, int min_auto_threshold, uint32_t* cutoff_out, uint64_t* nb_solids, uint32_t* first_peak)
{
    if (!h || histo_max < 1) return 1;
    const size_t length = histo_max;
    std::vector<uint64_t> sm (length + 2, 0);
    uint64_t sum_allk = 0; uint32_t cutoff = 0, peak = 0;
    if (length >= 2) { sm[1] = (uint64_t)(0.6 * (double)h[1] + 0.4 * (double)h[2]); sum_allk += h[1]; }
    int first_inc = -1, idx_max = -1; uint64_t max_val = 0;
    for (size_t i = 2; i < length; i++)
    {
        sum_allk += h[i] * i;
        sm[i] = (uint64_t)(0.2 * (double)h[i-1] + 0.6 * (double)h[i] + 0.2 * (double)h[i+1]);
        if (first_inc == -1 && sm[i-1] < sm[i]) first_inc = (int)i - 1;
        if (first_inc > 0 && sm[i] > max_val) { max_val = sm[i]; idx_max = (int)i; }
    }
    sum_allk += h[length] * length;
    if (first_inc == -1) { *cutoff_out = (uint32_t)min_auto_threshold; *nb_solids = 0; *first_peak = 0; return 0; }
    peak = (uint32_t)idx_max;
    uint64_t min_val = 10000000000ULL; int idx_min = -1;
    for (int i = first_inc; i <= idx_max; i++) if (sm[i] < min_val) { min_val = sm[i]; idx_min = i; }
    if (idx_min != -1) cutoff = (uint32_t)idx_min;
    uint64_t sum_elim = 0; size_t max_cutoff = 0;
    for (size_t i = 0; i < length + 1; i++)
    {
        sum_elim += h[i] * i;
        double ratio = (double)sum_elim / sum_allk;
        if (ratio >= 0.25) { max_cutoff = i + 1; break; }
    }
    if (cutoff > max_cutoff) cutoff = (uint32_t)max_cutoff;
    if (cutoff < (size_t)min_auto_threshold) cutoff = (uint32_t)min_auto_threshold;
    uint64_t nbs = 0; for (size_t i = cutoff; i < length + 1; i++) nbs += h[i];
    *cutoff_out = cutoff; *nb_solids = nbs; *first_peak = peak;
    return 0;
}

} // extern "C"
