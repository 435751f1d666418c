// pipe.cu — frame sink: pinned host ring + copy stream + writer thread.
// Replaces turbopipe.pipe/sync/done and fbo.read_into of ExportingHelper.pipe (exporting.py:140-174):
// the reference maps N GL buffers and lets a background thread write() them to ffmpeg's stdin; here
// the ring is cudaHostAlloc'd memory filled by an async D2H that is ordered after the render stream.
#include "sfb_internal.h"

#include <cerrno>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <unistd.h>
#include <vector>

struct sfb_pipe {
    sfb_ctx* ctx = nullptr;
    int fd = -1;
    size_t frame_bytes = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t rendered = nullptr;
    struct Slot { void* host = nullptr; void* dev = nullptr; cudaEvent_t copied = nullptr; int state = 0; /*0 free 1 queued*/ };
    bool acquired = false;
    std::vector<Slot> slots;
    size_t head = 0;                 // next slot to fill
    size_t tail = 0;                 // next slot to write
    uint64_t submitted = 0, written = 0, bytes = 0;
    bool closing = false;
    int io_errno = 0;
    std::mutex mu;
    std::condition_variable cv;
    std::thread writer;
};

static void writer_loop(sfb_pipe* p) {
    cudaSetDevice(p->ctx->device);
    for (;;) {
        sfb_pipe::Slot* slot = nullptr;
        {
            std::unique_lock<std::mutex> lock(p->mu);
            p->cv.wait(lock, [&] { return p->closing || p->slots[p->tail % p->slots.size()].state == 1; });
            if (p->slots[p->tail % p->slots.size()].state != 1) return;       // closing and drained
            slot = &p->slots[p->tail % p->slots.size()];
        }
        cudaEventSynchronize(slot->copied);
        int fd, failed;
        { std::lock_guard<std::mutex> lock(p->mu); fd = p->fd; failed = p->io_errno; }
        if (fd >= 0 && !failed) {
            const char* src = static_cast<const char*>(slot->host);
            size_t left = p->frame_bytes;
            while (left) {
                ssize_t n = ::write(fd, src, left);
                if (n < 0) { if (errno == EINTR) continue; failed = errno; break; }
                src += n; left -= size_t(n);
            }
        }
        {
            std::lock_guard<std::mutex> lock(p->mu);
            if (failed) p->io_errno = failed;
            slot->state = 0; p->tail++; p->written++; p->bytes += p->frame_bytes;
        }
        p->cv.notify_all();
    }
}

extern "C" int sfb_pipe_open(sfb_ctx* ctx, int fd, int n_buffers, size_t frame_bytes, sfb_pipe** out) {
    SFB_REQUIRE(ctx && out, "sfb_pipe_open: null argument");
    SFB_REQUIRE(n_buffers >= 1 && n_buffers <= 64 && frame_bytes > 0, "sfb_pipe_open: bad ring %d x %zu", n_buffers, frame_bytes);
    sfb_pipe* p = new sfb_pipe();
    p->ctx = ctx; p->fd = fd; p->frame_bytes = frame_bytes;
    p->slots.resize(size_t(n_buffers));
    cudaError_t e = cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->rendered, cudaEventDisableTiming);
    for (auto& s : p->slots) {
        // reuse a pooled (pinned host, device) pair of the right size when the context has one
        for (size_t i = 0; i < ctx->ring_pool.size() && !s.host; i++)
            if (ctx->ring_pool[i].bytes == frame_bytes) {
                s.host = ctx->ring_pool[i].host; s.dev = ctx->ring_pool[i].dev;
                ctx->ring_pool.erase(ctx->ring_pool.begin() + i);
            }
        if (!s.host) {
            if (e == cudaSuccess) e = cudaHostAlloc(&s.host, frame_bytes, cudaHostAllocDefault);
            if (e == cudaSuccess) e = cudaMalloc(&s.dev, frame_bytes);
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        for (auto& s : p->slots) { if (s.host) cudaFreeHost(s.host); if (s.dev) cudaFree(s.dev); if (s.copied) cudaEventDestroy(s.copied); }
        if (p->rendered) cudaEventDestroy(p->rendered);
        if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
        delete p;
        SFB_FAIL(SFB_ECUDA, "sfb_pipe_open: %s", cudaGetErrorString(e));
    }
    p->writer = std::thread(writer_loop, p);
    *out = p;
    return SFB_OK;
}

// Next device frame of the ring; blocks while its previous contents are still on their way out
extern "C" int sfb_pipe_acquire(sfb_pipe* p, void** frame_dev) {
    SFB_REQUIRE(p && frame_dev, "sfb_pipe_acquire: null argument");
    std::unique_lock<std::mutex> lock(p->mu);
    if (p->io_errno) SFB_FAIL(SFB_EIO, "sink write failed: %s", strerror(p->io_errno));
    sfb_pipe::Slot* slot = &p->slots[p->head % p->slots.size()];
    p->cv.wait(lock, [&] { return slot->state == 0; });             // ring full → wait for the writer
    p->acquired = true;
    *frame_dev = slot->dev;
    return SFB_OK;
}

extern "C" int sfb_pipe_submit(sfb_pipe* p, const void* frame_dev) {
    SFB_REQUIRE(p, "sfb_pipe_submit: null pipe");
    if (!p->acquired) { void* unused = nullptr; if (int e = sfb_pipe_acquire(p, &unused)) return e; }
    sfb_pipe::Slot* slot = &p->slots[p->head % p->slots.size()];
    if (frame_dev && frame_dev != slot->dev)                         // foreign buffer: stage it (D2D, render stream)
        SFB_CUDA(cudaMemcpyAsync(slot->dev, frame_dev, p->frame_bytes, cudaMemcpyDeviceToDevice, p->ctx->stream));
    // D2H ordered after everything enqueued so far on the render stream
    SFB_CUDA(cudaEventRecord(p->rendered, p->ctx->stream));
    SFB_CUDA(cudaStreamWaitEvent(p->copy_stream, p->rendered, 0));
    SFB_CUDA(cudaMemcpyAsync(slot->host, slot->dev, p->frame_bytes, cudaMemcpyDeviceToHost, p->copy_stream));
    SFB_CUDA(cudaEventRecord(slot->copied, p->copy_stream));
    {
        std::lock_guard<std::mutex> lock(p->mu);
        slot->state = 1; p->head++; p->submitted++; p->acquired = false;
    }
    p->cv.notify_all();
    return SFB_OK;
}

extern "C" int sfb_pipe_set_fd(sfb_pipe* p, int fd) {
    SFB_REQUIRE(p, "sfb_pipe_set_fd: null pipe");
    std::unique_lock<std::mutex> lock(p->mu);
    if (p->written != p->submitted || p->acquired)
        SFB_FAIL(SFB_ESTATE, "sfb_pipe_set_fd: the ring still has frames in flight; call sfb_pipe_sync first");
    p->fd = fd; p->io_errno = 0;
    return SFB_OK;
}

extern "C" int sfb_pipe_sync(sfb_pipe* p) {
    SFB_REQUIRE(p, "sfb_pipe_sync: null pipe");
    std::unique_lock<std::mutex> lock(p->mu);
    p->cv.wait(lock, [&] { return p->written == p->submitted; });
    if (p->io_errno) SFB_FAIL(SFB_EIO, "sink write failed: %s", strerror(p->io_errno));
    return SFB_OK;
}

extern "C" int sfb_pipe_stats(sfb_pipe* p, uint64_t* frames, uint64_t* bytes) {
    SFB_REQUIRE(p, "sfb_pipe_stats: null pipe");
    std::lock_guard<std::mutex> lock(p->mu);
    if (frames) *frames = p->written;
    if (bytes) *bytes = p->bytes;
    return SFB_OK;
}

extern "C" int sfb_pipe_close(sfb_pipe* p) {
    if (!p) return SFB_OK;
    {
        std::unique_lock<std::mutex> lock(p->mu);
        p->cv.wait(lock, [&] { return p->written == p->submitted; });
        p->closing = true;
    }
    p->cv.notify_all();
    p->writer.join();
    const int err = p->io_errno;
    cudaStreamSynchronize(p->copy_stream);
    for (auto& s : p->slots) {
        cudaEventDestroy(s.copied);
        if (p->ctx->ring_pool.size() < 16) p->ctx->ring_pool.push_back({s.host, s.dev, p->frame_bytes});
        else { cudaFreeHost(s.host); cudaFree(s.dev); }
    }
    cudaEventDestroy(p->rendered);
    cudaStreamDestroy(p->copy_stream);
    delete p;
    if (err) SFB_FAIL(SFB_EIO, "sink write failed: %s", strerror(err));
    return SFB_OK;
}
