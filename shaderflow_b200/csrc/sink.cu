// sink.cu — the sink of a frame-SHARDED export: every rank drains its own frames over its own PCIe link.
//
// The reference has one process, one GL context and one turbopipe writer (exporting.py:140-174). When the
// export is sharded over the GPUs of a box the rgb24 stream still has to reach ONE file
// descriptor in time order. Funnelling the frames into rank 0's HBM and then through rank 0's single PCIe
// link capped 8 GPUs at 1.8 k 4K frames/s (round 1: 45 GB/s, seven links idle). Here the reassembly happens
// in HOST memory instead:
//
//   * one shared-memory segment (memfd, or POSIX shm as a fallback) holds a ring of `slots` frames per rank;
//     each rank pins ITS ring (cudaHostRegister) and copies device → host into it on its own copy stream,
//     ordered after the render stream, as soon as a frame is shaded;
//   * ownership is block-cyclic — global frame g belongs to rank (g / block) % world — so all links are busy
//     at the same time and the rings stay small whatever the length of the export;
//   * rank 0's writer thread walks the frames in time order, waits for the owner's `produced` counter, write()s
//     the frame from the owner's ring to the sink's fd (or drops it: null sink) and bumps `consumed`, which is
//     the back-pressure the owners block on in sfb_sink_acquire (the role of turbopipe.sync, exporting.py:168).
//
// The counters live in the segment header, one cache line per rank and direction; they are the only cross-
// process communication (no collective, no kernel, no SM). ctx == NULL opens the sink in host-only mode: frames
// are handed over as host pointers (sfb_sink_submit_host) — the `-m "not gpu"` tests drive the same protocol
// over gloo-launched processes that way.
#include "sfb_internal.h"

#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstring>
#include <fcntl.h>
#include <new>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

namespace {

constexpr uint32_t SINK_MAGIC = 0x53464253u;      // "SFBS"
constexpr int SINK_MAX_WORLD = 64;
constexpr size_t SINK_PAGE = 4096;

struct alignas(128) RankCounters {
    std::atomic<uint64_t> produced;               // frames of this rank that are complete in its ring (owner writes)
    char pad0[128 - sizeof(std::atomic<uint64_t>)];
    std::atomic<uint64_t> consumed;               // frames of this rank the writer is done with (rank 0 writes)
    char pad1[128 - sizeof(std::atomic<uint64_t>)];
};

struct SinkHeader {
    uint32_t magic, version;
    int32_t world, slots;
    uint64_t frame_bytes, frame_stride;
    std::atomic<int32_t> abort;                   // non-zero: an error somewhere, everybody unblocks
    std::atomic<int32_t> io_errno;
    RankCounters rank[SINK_MAX_WORLD];
};

size_t header_bytes() { return (sizeof(SinkHeader) + SINK_PAGE - 1)/SINK_PAGE*SINK_PAGE; }

}  // namespace

struct sfb_sink {
    sfb_ctx* ctx = nullptr;
    int rank = 0, world = 1, slots = 0;
    size_t frame_bytes = 0, frame_stride = 0, segment_bytes = 0;
    unsigned char* base = nullptr;                // the mapping
    SinkHeader* header = nullptr;
    int mem_fd = -1;
    std::string path;
    bool registered = false;
    // export state
    int64_t total_frames = 0;
    int block = 1;
    uint64_t mine = 0;                            // frames this rank owns in the current export
    uint64_t next = 0;                            // local sequence number of the next frame to acquire
    bool acquired = false;
    double timeout_s = 300.0;
    // device side
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t rendered = nullptr;
    std::vector<void*> dev;                       // one device frame per slot
    // rank 0
    std::thread writer;
    bool writer_running = false;
    int fd = -1;
    uint64_t written = 0, bytes = 0;

    unsigned char* ring(int r) const { return base + header_bytes() + size_t(r)*size_t(slots)*frame_stride; }
    unsigned char* slot(int r, uint64_t q) const { return ring(r) + size_t(q % uint64_t(slots))*frame_stride; }
};

// Block-cyclic ownership: frame g → (owner rank, sequence number among the owner's frames)
static inline void owner_of(int64_t g, int block, int world, int* owner, uint64_t* seq) {
    const int64_t b = g/block;
    *owner = int(b % world);
    *seq = uint64_t((b/world)*block + (g % block));
}

static uint64_t frames_owned(int64_t total, int block, int world, int rank) {
    uint64_t n = 0;
    for (int64_t b = rank; b*block < total; b += world) {
        const int64_t left = total - b*block;
        n += uint64_t(left < block ? left : block);
    }
    return n;
}

template <typename Pred> static bool wait_until(sfb_sink* s, Pred done) {
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spins = 0; !done(); spins++) {
        if (s->header->abort.load(std::memory_order_acquire)) return false;
        if (spins < 200) { std::this_thread::yield(); continue; }
        std::this_thread::sleep_for(std::chrono::microseconds(spins < 2000 ? 20 : 200));
        if ((spins & 1023) == 0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > s->timeout_s) {
            s->header->abort.store(2, std::memory_order_release);
            return false;
        }
    }
    return true;
}

static int sink_failed(sfb_sink* s, const char* who) {
    const int err = s->header->io_errno.load();
    if (err) SFB_FAIL(SFB_EIO, "%s: sink write failed: %s", who, strerror(err));
    if (s->header->abort.load() == 2) SFB_FAIL(SFB_ESTATE, "%s: timed out waiting for another rank of the sharded export", who);
    SFB_FAIL(SFB_ESTATE, "%s: the sharded export was aborted by another rank", who);
}

static void writer_loop(sfb_sink* s) {
    SinkHeader* h = s->header;
    for (int64_t g = 0; g < s->total_frames; g++) {
        int owner; uint64_t q;
        owner_of(g, s->block, s->world, &owner, &q);
        if (!wait_until(s, [&] { return h->rank[owner].produced.load(std::memory_order_acquire) > q; })) return;
        if (s->fd >= 0) {
            const unsigned char* src = s->slot(owner, q);
            size_t left = s->frame_bytes;
            while (left) {
                const ssize_t n = ::write(s->fd, src, left);
                if (n < 0) {
                    if (errno == EINTR) continue;
                    h->io_errno.store(errno); h->abort.store(1, std::memory_order_release);
                    return;
                }
                src += n; left -= size_t(n);
            }
        }
        h->rank[owner].consumed.store(q + 1, std::memory_order_release);
        s->written++; s->bytes += s->frame_bytes;
    }
}

static void sink_destroy(sfb_sink* s) {
    if (s->writer_running) { s->header->abort.store(1); s->writer.join(); s->writer_running = false; }
    if (s->ctx && s->base) {
        cudaSetDevice(s->ctx->device);
        if (s->copy_stream) cudaStreamSynchronize(s->copy_stream);
        if (s->registered) cudaHostUnregister(s->ring(s->rank));
        for (void* d : s->dev) if (d) cudaFree(d);
        if (s->rendered) cudaEventDestroy(s->rendered);
        if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    }
    if (s->base) munmap(s->base, s->segment_bytes);
    if (s->mem_fd >= 0) close(s->mem_fd);
    if (s->rank == 0 && s->path.rfind("/dev/shm/", 0) == 0) unlink(s->path.c_str());
    delete s;
}

extern "C" int sfb_sink_open(sfb_ctx* ctx, const char* segment_path, int rank, int world, int slots,
                             size_t frame_bytes, sfb_sink** out) {
    SFB_REQUIRE(out, "sfb_sink_open: null argument");
    SFB_REQUIRE(world >= 1 && world <= SINK_MAX_WORLD && rank >= 0 && rank < world, "sfb_sink_open: rank %d of %d", rank, world);
    SFB_REQUIRE(slots >= 2 && slots <= 256 && frame_bytes > 0, "sfb_sink_open: bad ring %d x %zu", slots, frame_bytes);
    SFB_REQUIRE((rank == 0) == (segment_path == nullptr || segment_path[0] == 0),
        "sfb_sink_open: rank 0 creates the segment (no path), every other rank attaches to rank 0's path");
    sfb_sink* s = new sfb_sink();
    s->ctx = ctx; s->rank = rank; s->world = world; s->slots = slots; s->frame_bytes = frame_bytes;
    s->frame_stride = (frame_bytes + SINK_PAGE - 1)/SINK_PAGE*SINK_PAGE;
    s->segment_bytes = header_bytes() + size_t(world)*size_t(slots)*s->frame_stride;
    if (const char* t = getenv("SFB_SINK_TIMEOUT_S")) s->timeout_s = atof(t);
    if (rank == 0) {
        static std::atomic<int> serial{0};
        char name[96];
        snprintf(name, sizeof(name), "sfb200-sink-%d-%d", int(getpid()), serial.fetch_add(1));
        s->mem_fd = memfd_create(name, MFD_CLOEXEC);
        if (s->mem_fd >= 0) {
            snprintf(name, sizeof(name), "/proc/%d/fd/%d", int(getpid()), s->mem_fd);
            s->path = name;
        } else {                                  // no memfd: POSIX shared memory
            std::string shm = std::string("/") + name;
            s->mem_fd = shm_open(shm.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
            s->path = "/dev/shm" + shm;
        }
        if (s->mem_fd < 0 || ftruncate(s->mem_fd, off_t(s->segment_bytes)) != 0
            || posix_fallocate(s->mem_fd, 0, off_t(s->segment_bytes)) != 0) {
            const int err = errno; const size_t want = s->segment_bytes; sink_destroy(s);
            SFB_FAIL(SFB_ENOMEM, "sfb_sink_open: cannot create a %zu-byte shared segment: %s", want, strerror(err));
        }
    } else {
        s->path = segment_path;
        s->mem_fd = open(segment_path, O_RDWR | O_CLOEXEC);
        struct stat st;
        if (s->mem_fd < 0 || fstat(s->mem_fd, &st) != 0 || size_t(st.st_size) != s->segment_bytes) {
            const int err = errno; sink_destroy(s);
            SFB_FAIL(SFB_ESTATE, "sfb_sink_open: cannot attach to segment '%s' (%s, or its size differs)", segment_path, strerror(err));
        }
    }
    void* map = mmap(nullptr, s->segment_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, s->mem_fd, 0);
    if (map == MAP_FAILED) { const int err = errno; sink_destroy(s); SFB_FAIL(SFB_ENOMEM, "sfb_sink_open: mmap: %s", strerror(err)); }
    s->base = static_cast<unsigned char*>(map);
    s->header = reinterpret_cast<SinkHeader*>(s->base);
    if (rank == 0) {
        new (s->header) SinkHeader();
        s->header->version = 1; s->header->world = world; s->header->slots = slots;
        s->header->frame_bytes = frame_bytes; s->header->frame_stride = s->frame_stride;
        s->header->abort.store(0); s->header->io_errno.store(0);
        for (int r = 0; r < SINK_MAX_WORLD; r++) { s->header->rank[r].produced.store(0); s->header->rank[r].consumed.store(0); }
        std::atomic_thread_fence(std::memory_order_seq_cst);
        s->header->magic = SINK_MAGIC;
    } else if (s->header->magic != SINK_MAGIC || s->header->world != world || s->header->slots != slots
               || s->header->frame_bytes != frame_bytes) {
        sink_destroy(s);
        SFB_FAIL(SFB_ESTATE, "sfb_sink_open: segment '%s' was made for another ring geometry", segment_path);
    }
    if (ctx) {
        cudaError_t e = cudaSetDevice(ctx->device);
        // pin THIS rank's ring: the D2H copies go straight into the shared pages
        if (e == cudaSuccess) e = cudaHostRegister(s->ring(rank), size_t(slots)*s->frame_stride, cudaHostRegisterPortable);
        s->registered = (e == cudaSuccess);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->rendered, cudaEventDisableTiming);
        s->dev.assign(size_t(slots), nullptr);
        for (int k = 0; k < slots && e == cudaSuccess; k++) e = cudaMalloc(&s->dev[size_t(k)], frame_bytes);
        if (e != cudaSuccess) {
            sink_destroy(s);
            SFB_FAIL(SFB_ECUDA, "sfb_sink_open: %s", cudaGetErrorString(e));
        }
    }
    *out = s;
    return SFB_OK;
}

extern "C" const char* sfb_sink_path(sfb_sink* s) { return s ? s->path.c_str() : ""; }

// Starts one export of `total_frames` frames, block-cyclic with `block` frames per block. Collective by
// convention: every rank calls it, then the caller synchronises the ranks (a barrier) BEFORE any frame is
// produced — rank 0 resets the shared counters here. fd: rank 0's sink (-1 = null sink), ignored elsewhere.
extern "C" int sfb_sink_begin(sfb_sink* s, int64_t total_frames, int block, int fd) {
    SFB_REQUIRE(s, "sfb_sink_begin: null sink");
    SFB_REQUIRE(total_frames >= 0 && block >= 1 && block <= s->slots, "sfb_sink_begin: block %d must be 1..slots (%d)", block, s->slots);
    SFB_REQUIRE(!s->writer_running && !s->acquired, "sfb_sink_begin: the previous export is still running");
    s->total_frames = total_frames; s->block = block;
    s->mine = frames_owned(total_frames, block, s->world, s->rank);
    s->next = 0; s->written = 0; s->bytes = 0; s->fd = fd;
    if (s->rank == 0) {
        s->header->abort.store(0); s->header->io_errno.store(0);
        for (int r = 0; r < s->world; r++) { s->header->rank[r].produced.store(0); s->header->rank[r].consumed.store(0); }
        std::atomic_thread_fence(std::memory_order_seq_cst);
        s->writer = std::thread(writer_loop, s);
        s->writer_running = true;
    }
    return SFB_OK;
}

extern "C" int sfb_sink_owner(sfb_sink* s, int64_t frame, int* owner) {
    SFB_REQUIRE(s && owner && frame >= 0, "sfb_sink_owner: bad argument");
    uint64_t q; owner_of(frame, s->block, s->world, owner, &q);
    return SFB_OK;
}

// Device frame the next frame this rank owns is rendered into; blocks while the writer has not consumed the
// frame that used the slot `slots` frames ago (back-pressure)
extern "C" int sfb_sink_acquire(sfb_sink* s, void** frame_dev) {
    SFB_REQUIRE(s && frame_dev, "sfb_sink_acquire: null argument");
    SFB_REQUIRE(s->next < s->mine, "sfb_sink_acquire: this rank owns %llu frames of the export and has produced them all", (unsigned long long)s->mine);
    const uint64_t q = s->next;
    SinkHeader* h = s->header;
    if (!wait_until(s, [&] { return h->rank[s->rank].consumed.load(std::memory_order_acquire) + uint64_t(s->slots) > q; }))
        return sink_failed(s, "sfb_sink_acquire");
    s->acquired = true;
    *frame_dev = s->ctx ? s->dev[size_t(q % uint64_t(s->slots))] : static_cast<void*>(s->slot(s->rank, q));
    return SFB_OK;
}

struct Publish { std::atomic<uint64_t>* counter; uint64_t value; };
static void CUDART_CB publish_cb(void* arg) {
    Publish* p = static_cast<Publish*>(arg);
    p->counter->store(p->value, std::memory_order_release);
    delete p;
}

// The frame rendered into the acquired device frame: D2H into this rank's ring (ordered after the render
// stream), then `produced` is published from the copy stream
extern "C" int sfb_sink_submit(sfb_sink* s) {
    SFB_REQUIRE(s && s->ctx, "sfb_sink_submit: needs a device context (host-only sinks use sfb_sink_submit_host)");
    SFB_REQUIRE(s->acquired, "sfb_sink_submit: no frame acquired");
    const uint64_t q = s->next;
    SFB_CUDA(cudaEventRecord(s->rendered, s->ctx->stream));
    SFB_CUDA(cudaStreamWaitEvent(s->copy_stream, s->rendered, 0));
    SFB_CUDA(cudaMemcpyAsync(s->slot(s->rank, q), s->dev[size_t(q % uint64_t(s->slots))], s->frame_bytes, cudaMemcpyDeviceToHost, s->copy_stream));
    SFB_CUDA(cudaLaunchHostFunc(s->copy_stream, publish_cb, new Publish{&s->header->rank[s->rank].produced, q + 1}));
    s->next = q + 1; s->acquired = false;
    return SFB_OK;
}

// Host-only hand-over (tests, CPU producers): frame_host NULL = the acquired slot was filled in place
extern "C" int sfb_sink_submit_host(sfb_sink* s, const void* frame_host) {
    SFB_REQUIRE(s, "sfb_sink_submit_host: null sink");
    if (!s->acquired) { void* unused; if (int e = sfb_sink_acquire(s, &unused)) return e; }
    const uint64_t q = s->next;
    if (s->ctx) SFB_CUDA(cudaStreamSynchronize(s->copy_stream));
    if (frame_host && frame_host != s->slot(s->rank, q)) memcpy(s->slot(s->rank, q), frame_host, s->frame_bytes);
    s->header->rank[s->rank].produced.store(q + 1, std::memory_order_release);
    s->next = q + 1; s->acquired = false;
    return SFB_OK;
}

// Blocks until every frame this rank owns has left through the writer (rank 0: until the writer wrote the
// whole export). frames / bytes: what rank 0's writer wrote (0 elsewhere).
extern "C" int sfb_sink_finish(sfb_sink* s, uint64_t* frames, uint64_t* bytes) {
    SFB_REQUIRE(s, "sfb_sink_finish: null sink");
    SFB_REQUIRE(s->next == s->mine, "sfb_sink_finish: this rank produced %llu of its %llu frames", (unsigned long long)s->next, (unsigned long long)s->mine);
    SinkHeader* h = s->header;
    bool ok = wait_until(s, [&] { return h->rank[s->rank].consumed.load(std::memory_order_acquire) >= s->mine; });
    if (s->writer_running) { s->writer.join(); s->writer_running = false; ok = ok && s->written == uint64_t(s->total_frames); }
    if (frames) *frames = s->written;
    if (bytes) *bytes = s->bytes;
    if (!ok) return sink_failed(s, "sfb_sink_finish");
    return SFB_OK;
}

// Unblocks every rank of the export with an error (a rank that fails calls this before it leaves)
extern "C" int sfb_sink_abort(sfb_sink* s) {
    if (s && s->header) s->header->abort.store(1, std::memory_order_release);
    if (s && s->writer_running) { s->writer.join(); s->writer_running = false; }
    if (s) { s->acquired = false; s->next = s->mine; }
    return SFB_OK;
}

extern "C" int sfb_sink_close(sfb_sink* s) {
    if (s) sink_destroy(s);
    return SFB_OK;
}
