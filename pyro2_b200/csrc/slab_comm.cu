// slab_comm.cu -- halo rows of the STATE planes of a decomposed run through peer memory (HP-1 and the explicit solvers).
//
// The reference is single-process (pyro/mesh/array_indexer.py:157-158: "there is only a single grid"); this is the B200-side
// extension SURVEY.md 8(e) describes.  Each rank owns an x-slab; before a step its ng ghost rows on a side that faces another
// slab must hold that slab's boundary rows (what ArrayIndexer.fill_ghost's periodic / interior copy is to one grid), and
// the time step needs the maximum wave speeds over all slabs.  Both used to be NCCL calls (16 point-to-point messages, an
// all-reduce and a host sync: 0.45 ms of a 2.4 ms step).  Here a rank stores its boundary rows STRAIGHT INTO the
// neighbours' ghost rows (the planes are cudaMalloc'd by the library and mapped into every rank, p2b_shared_*), and the
// maxima are combined through per-rank slots on every rank; ordering by monotone 64-bit words with bounded spins, as in
// the multigrid (peer_comm.cuh, mg_kernels.cuh).  Three tiny launches per exchange, one per reduction, no host involvement.
//
// Protocol of one exchange (a "program"; all ranks call the same sequence):
//   arrive  <<<1,1>>>  epoch += 1; tell both neighbours "everything I enqueued before has finished: my ghost rows may be
//                      overwritten" (value 2 epoch); wait for the same from them
//   push    grid       copy my first / last ng valid rows of every plane into the neighbours' ghost rows; the last CTA
//                      (local counter after a system-scope fence) stores 2 epoch + 1 into both neighbours' words
//   wait    <<<1,32>>> wait for both neighbours' 2 epoch + 1
#include "common.cuh"
#include "peer_comm.cuh"

namespace pyro {

constexpr int SLAB_MAX_RANKS = 16;
constexpr int SLAB_MAX_BUFFERS = 64;    // plane buffers one communicator can hold (two per compressible simulation)
enum { SW_EPOCH = 0, SW_ERR = COMM_ERR_WORD, SW_FROM_LO = 2, SW_FROM_HI = 3, SW_CNT = 4, SW_GFLAG = 8 /* [16] */,
       SW_SLOTS = 24 /* u64 [2][16][4] */, SW_WORDS = 24 + 2 * 16 * 4 };

struct SlabComm {
    unsigned long long* ctl;                 // this rank's control words
    long long clo, chi;                      // element offsets (u64) to the lo / hi neighbour's control words
    int has_lo, has_hi, rank, size;
};

__global__ void slab_arrive_kernel(SlabComm c)
{
    c.ctl[SW_EPOCH] += 1;
    const unsigned long long val = 2ull * c.ctl[SW_EPOCH];
    __threadfence_system();
    if (c.has_lo) comm_store(c.ctl + c.clo + SW_FROM_HI, val);
    if (c.has_hi) comm_store(c.ctl + c.chi + SW_FROM_LO, val);
    if (c.has_lo) comm_wait_ge(c.ctl, SW_FROM_LO, val);
    if (c.has_hi) comm_wait_ge(c.ctl, SW_FROM_HI, val);
}

// grid (ceil(pitch / 256), ng, nvar); dlo / dhi: element offsets from my planes to the lo / hi neighbour's
__global__ void slab_push_kernel(double* planes, long long plane_stride, int pitch, int nx, int ng, long long dlo, long long dhi,
                                 SlabComm c)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    double* p = planes + (long long)blockIdx.z * plane_stride;
    if (col < pitch) {
        // my bottom valid rows ng .. 2 ng - 1 are the lo neighbour's high ghost rows ng + nx ..; my top valid rows
        // nx .. nx + ng - 1 are the hi neighbour's low ghost rows 0 .. ng - 1
        if (c.has_lo) (p + dlo)[(long long)(ng + nx + r) * pitch + col] = p[(long long)(ng + r) * pitch + col];
        if (c.has_hi) (p + dhi)[(long long)r * pitch + col] = p[(long long)(nx + r) * pitch + col];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned long long nblocks = (unsigned long long)gridDim.x * gridDim.y * gridDim.z;
        if (atomicAdd(c.ctl + SW_CNT, 1ull) == nblocks - 1) {
            atomicExch(c.ctl + SW_CNT, 0ull);
            const unsigned long long val = 2ull * c.ctl[SW_EPOCH] + 1ull;
            if (c.has_lo) comm_store(c.ctl + c.clo + SW_FROM_HI, val);
            if (c.has_hi) comm_store(c.ctl + c.chi + SW_FROM_LO, val);
        }
    }
}

__global__ void slab_wait_kernel(SlabComm c)
{
    const unsigned long long val = 2ull * c.ctl[SW_EPOCH] + 1ull;
    if (threadIdx.x == 0 && c.has_lo) comm_wait_ge(c.ctl, SW_FROM_LO, val);
    if (threadIdx.x == 1 && c.has_hi) comm_wait_ge(c.ctl, SW_FROM_HI, val);
}

// <<<1, 32>>>: words[0..3] <- max over the ranks (unsigned 64-bit: positive doubles order like their bit patterns).  Each
// rank writes its four words into its slot on every rank, publishes, waits for everybody's slot.  A program of its own;
// the slots are double-buffered by epoch parity (a fast rank's next reduction must not overwrite what a slow one reads).
__global__ void slab_allreduce_max4_kernel(unsigned long long* words, SlabComm c, const long long* peer_ctl_off)
{
    __shared__ unsigned long long mine[4];
    const int t = threadIdx.x;
    if (t == 0) c.ctl[SW_EPOCH] += 1;
    if (t < 4) mine[t] = words[t];
    __syncwarp();
    const unsigned long long ep = c.ctl[SW_EPOCH], val = 2ull * ep + 1ull;
    unsigned long long* slots = c.ctl + SW_SLOTS + (ep & 1ull) * SLAB_MAX_RANKS * 4;
    if (t < c.size) {
        unsigned long long* dst = slots + peer_ctl_off[t] + c.rank * 4;
        dst[0] = mine[0]; dst[1] = mine[1]; dst[2] = mine[2]; dst[3] = mine[3];
        __threadfence_system();
        comm_store_relaxed(c.ctl + peer_ctl_off[t] + SW_GFLAG + c.rank, val);
        comm_wait_ge(c.ctl, SW_GFLAG + t, val);
    }
    __syncwarp();
    if (t < 4) {
        unsigned long long m = 0;
        for (int r = 0; r < c.size; ++r) {
            const volatile unsigned long long* src = slots + r * 4;
            m = src[t] > m ? src[t] : m;
        }
        words[t] = m;
    }
}

}  // namespace pyro

using namespace pyro;

struct p2b_slab {
    int rank, size, periodic;
    unsigned long long* ctl;
    unsigned long long* peer_ctl[SLAB_MAX_RANKS];
    long long* peer_off_dev;                        // device copy of the control-word offsets (inside the ctl block's tail)
    int nbuf;
    double* base[SLAB_MAX_BUFFERS];
    long long bytes[SLAB_MAX_BUFFERS];
    double* peer_base[SLAB_MAX_BUFFERS][SLAB_MAX_RANKS];
};

static SlabComm slab_comm(const p2b_slab* s)
{
    SlabComm c;
    c.ctl = s->ctl; c.rank = s->rank; c.size = s->size;
    c.has_lo = (s->rank > 0 || s->periodic) ? 1 : 0;
    c.has_hi = (s->rank < s->size - 1 || s->periodic) ? 1 : 0;
    if (s->size == 1) c.has_lo = c.has_hi = 0;
    const int lo = (s->rank + s->size - 1) % s->size, hi = (s->rank + 1) % s->size;
    c.clo = s->peer_ctl[lo] - s->ctl;
    c.chi = s->peer_ctl[hi] - s->ctl;
    return c;
}

extern "C" {

// bytes of the control block every rank allocates with p2b_shared_alloc (zeroed) and shares
long long p2b_slab_ctl_bytes(void) { return (long long)(SW_WORDS + SLAB_MAX_RANKS) * 8; }

// rank `rank` of `size` x-slabs (periodic: the x direction wraps); ctl_peers[r] = rank r's control block as mapped here
p2b_slab* p2b_slab_create(int rank, int size, int periodic, void* const* ctl_peers)
{
    if (size < 1 || size > SLAB_MAX_RANKS || rank < 0 || rank >= size || !ctl_peers) { set_error("bad slab communicator arguments"); return nullptr; }
    p2b_slab* s = new p2b_slab();
    memset(s, 0, sizeof *s);
    s->rank = rank; s->size = size; s->periodic = periodic;
    for (int r = 0; r < size; ++r) s->peer_ctl[r] = (unsigned long long*)ctl_peers[r];
    s->ctl = s->peer_ctl[rank];
    long long off[SLAB_MAX_RANKS] = {0};
    for (int r = 0; r < size; ++r) off[r] = s->peer_ctl[r] - s->ctl;
    s->peer_off_dev = (long long*)(s->ctl + SW_WORDS);
    if (cudaMemcpy(s->peer_off_dev, off, sizeof off, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("cudaMemcpy failed"); delete s; return nullptr; }
#ifndef P2B_EMU_HEADER
    cudaFuncAttributes a;      // load now: a lazily loaded kernel must not meet a neighbour's spinning one (mg.cu, preload_kernels)
    cudaFuncGetAttributes(&a, slab_arrive_kernel); cudaFuncGetAttributes(&a, slab_push_kernel);
    cudaFuncGetAttributes(&a, slab_wait_kernel); cudaFuncGetAttributes(&a, slab_allreduce_max4_kernel);
    cudaGetLastError();
#endif
    return s;
}

int p2b_slab_destroy(p2b_slab* s) { delete s; return P2B_OK; }

// buffer k of this rank (`bytes` long, p2b_shared_alloc'd) and every rank's buffer k as mapped here: any plane set inside
// it can then be exchanged
int p2b_slab_register(p2b_slab* s, int k, void* local_base, long long bytes, void* const* peer_bases)
{
    P2B_REQUIRE(s && k >= 0 && k < SLAB_MAX_BUFFERS && local_base && peer_bases && bytes > 0, "bad buffer registration");
    P2B_REQUIRE(peer_bases[s->rank] == local_base, "peer_bases[rank] must be the local buffer");
    s->base[k] = (double*)local_base; s->bytes[k] = bytes;
    for (int r = 0; r < s->size; ++r) s->peer_base[k][r] = (double*)peer_bases[r];
    if (k >= s->nbuf) s->nbuf = k + 1;
    return P2B_OK;
}

// is [planes, planes + extent) inside a registered buffer?  (the caller falls back to its other transport otherwise)
int p2b_slab_owns(p2b_slab* s, const void* planes)
{
    if (!s) return 0;
    for (int k = 0; k < s->nbuf; ++k)
        if (s->base[k] && (const char*)planes >= (const char*)s->base[k] && (const char*)planes < (const char*)s->base[k] + s->bytes[k]) return 1;
    return 0;
}

// fill the ghost rows that face another slab with that slab's boundary rows: nvar planes (plane_stride apart) of
// (nx + 2 ng) rows x pitch doubles, same geometry on every rank.  Collective, asynchronous on `stream`.
int p2b_slab_exchange(p2b_slab* s, double* planes, int nvar, long long plane_stride, int pitch, int nx, int ng, void* stream)
{
    P2B_REQUIRE(s && planes && nvar >= 1 && pitch >= 1 && nx >= ng && ng >= 1, "bad exchange arguments");
    if (s->size == 1) return P2B_OK;
    int k = -1;
    for (int b = 0; b < s->nbuf; ++b)
        if (s->base[b] && planes >= s->base[b] && (const char*)planes < (const char*)s->base[b] + s->bytes[b]) k = b;
    P2B_REQUIRE(k >= 0, "planes are not inside a registered buffer");
    cudaStream_t st = (cudaStream_t)stream;
    SlabComm c = slab_comm(s);
    const int lo = (s->rank + s->size - 1) % s->size, hi = (s->rank + 1) % s->size;
    const long long dlo = s->peer_base[k][lo] - s->base[k], dhi = s->peer_base[k][hi] - s->base[k];
#ifdef P2B_EMU_HEADER
    ::emu::force_threaded = true;
#endif
    P2B_LAUNCH(slab_arrive_kernel, 1, 1, 0, st)(c);
    dim3 grd((pitch + 255) / 256, ng, nvar);
    P2B_LAUNCH(slab_push_kernel, grd, 256, 0, st)(planes, plane_stride, pitch, nx, ng, dlo, dhi, c);
    P2B_LAUNCH(slab_wait_kernel, 1, 32, 0, st)(c);
#ifdef P2B_EMU_HEADER
    ::emu::force_threaded = false;
#endif
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// words[0..3] (device, unsigned 64-bit) <- element-wise maximum over the ranks.  Collective, asynchronous.
int p2b_slab_allreduce_max4(p2b_slab* s, uint64_t* words, void* stream)
{
    P2B_REQUIRE(s && words, "null pointer");
    if (s->size == 1) return P2B_OK;
#ifdef P2B_EMU_HEADER
    ::emu::force_threaded = true;
#endif
    P2B_LAUNCH(slab_allreduce_max4_kernel, 1, 32, 0, (cudaStream_t)stream)((unsigned long long*)words, slab_comm(s), s->peer_off_dev);
#ifdef P2B_EMU_HEADER
    ::emu::force_threaded = false;
#endif
    P2B_CUDA_CHECK(cudaGetLastError());
    return P2B_OK;
}

// non-zero: a wait on another rank timed out (synchronises the stream)
int p2b_slab_error(p2b_slab* s, void* stream)
{
    if (!s) return 0;
    unsigned long long e = 0;
    if (cudaMemcpyAsync(&e, s->ctl + SW_ERR, 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess) return -1;
    cudaStreamSynchronize((cudaStream_t)stream);
    return (int)e;
}

}  // extern "C"
