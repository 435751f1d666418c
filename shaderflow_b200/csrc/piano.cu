// piano.cu — ShaderPiano.update (reference shaderflow/piano/module.py:202-277) on the GPU.
//
// The reference walks a Python dict-of-dict-of-deque note tree for every MIDI pitch on every frame and
// uploads three textures. Here the notes live in HBM once (sorted by pitch), and
//   sfb_piano_track  computes, for ALL frames of an export in one launch, what update() derives from the
//                    tree besides the roll: the key-press targets, the channel row and the range of
//                    upcoming pitches (one thread per (frame, pitch));
//   sfb_piano_roll   fills the (MAX_NOTE x MAX_ROLLING) roll texture of ONE frame straight into the
//                    texture's storage (one CTA per pitch, slots in the reference's iteration order);
//   sfb_dynamics_scan runs DynamicNumber.next over frames for a float32 vector (the 128 key presses).
//
// Iteration order of the reference (what decides the roll slots and which note "wins" a key):
// notes_between(pitch, t, t + lookup) visits the integer-second buckets int(t) .. int(t + lookup) in order
// and, inside a bucket, the notes in insertion order, each note once, at the first bucket it is found in:
// the key of a note is (max(int(start), int(t)), insertion index). A note is visited iff its buckets
// int(start) .. int(end) meet that range and start <= t + lookup (module.py:137-147).
#include "sfb_internal.h"

constexpr int PIANO_MAX_NOTE = 128, PIANO_MAX_ROLLING = 256, PIANO_CANDIDATES = 2048;

struct Visit { bool visited; long long bucket; };

// notes_between's membership test and the bucket a note is first met in
__device__ __forceinline__ Visit piano_visit(const sfb_piano_note& n, double t, double t_end) {
    const long long lo = (long long)t, hi = (long long)t_end;            // int(): truncation toward zero
    const long long n0 = (long long)n.start, n1 = (long long)n.end;
    Visit v;
    v.visited = (n1 >= lo) && (n0 <= hi) && (n1 >= n0) && !(n.start > t_end);
    v.bucket = n0 > lo ? n0 : lo;
    return v;
}

// (frame, pitch): key-press target, channel, "has upcoming notes" (module.py:213-247,262-266)
__global__ void __launch_bounds__(PIANO_MAX_NOTE) piano_track_kernel(const sfb_piano_note* notes, const int* pitch_offset,
        const double* time, int n_frames, double time_offset, double roll_time, double lookup_time, double release,
        int gmin, int gmax, float* key_target, float* channel, int* upcoming) {
    const int k = blockIdx.x, p = threadIdx.x;
    __shared__ int lo_s, hi_s;
    if (p == 0) { lo_s = PIANO_MAX_NOTE; hi_s = -1; }
    __syncthreads();
    float target = 0.0f, chan = -1.0f;
    bool any = false;
    if (p >= gmin && p <= gmax) {
        const double t = time[k] + time_offset, t_end = t + lookup_time;
        long long key_b = -1, chan_b = -1; int key_o = -1, chan_o = -1;
        for (int i = pitch_offset[p]; i < pitch_offset[p + 1]; i++) {
            const sfb_piano_note n = notes[i];
            const Visit v = piano_visit(n, t, t_end);
            if (!v.visited) continue;
            any = true;
            if (n.start >= t + roll_time) continue;
            if (!(n.start <= t && t <= n.end)) continue;
            // "last visited wins": keep the playing note with the largest (bucket, insertion index)
            const bool too_small = (n.end - n.start) < release;
            const bool shorter = t < (n.end - release);
            if ((shorter || too_small) && (v.bucket > key_b || (v.bucket == key_b && n.order > key_o))) {
                key_b = v.bucket; key_o = n.order; target = float(n.velocity);
            }
            if (v.bucket > chan_b || (v.bucket == chan_b && n.order > chan_o)) {
                chan_b = v.bucket; chan_o = n.order; chan = float(n.channel);
            }
        }
    }
    key_target[size_t(k)*PIANO_MAX_NOTE + p] = target;
    channel[size_t(k)*PIANO_MAX_NOTE + p] = chan;
    if (any) { atomicMin(&lo_s, p); atomicMax(&hi_s, p); }
    __syncthreads();
    if (p == 0) { upcoming[2*k] = lo_s; upcoming[2*k + 1] = hi_s; }        // (128, -1) when nothing is upcoming
}

// One frame's roll texture: CTA = pitch, texel (slot, pitch) = (start, end, channel, velocity) of the
// slot-th visited note that is inside the viewport (module.py:222-231); the other slots are zero
__global__ void __launch_bounds__(256) piano_roll_kernel(const sfb_piano_note* notes, const int* pitch_offset,
        double t, double roll_time, double lookup_time, int gmin, int gmax, float4* roll, int* overflow) {
    const int p = blockIdx.x, tid = threadIdx.x;
    __shared__ long long bucket[PIANO_CANDIDATES];
    __shared__ int index[PIANO_CANDIDATES];
    __shared__ int count;
    if (tid == 0) count = 0;
    for (int s = tid; s < PIANO_MAX_ROLLING; s += blockDim.x) roll[size_t(p)*PIANO_MAX_ROLLING + s] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    if (p < gmin || p > gmax) return;
    const double t_end = t + lookup_time;
    for (int i = pitch_offset[p] + tid; i < pitch_offset[p + 1]; i += blockDim.x) {
        const sfb_piano_note n = notes[i];
        const Visit v = piano_visit(n, t, t_end);
        if (v.visited && !(n.start >= t + roll_time)) {
            const int c = atomicAdd(&count, 1);
            if (c < PIANO_CANDIDATES) { bucket[c] = v.bucket; index[c] = i; }
        }
    }
    __syncthreads();
    int c = count;
    if (c > PIANO_CANDIDATES) { if (tid == 0) atomicExch(overflow, 1); c = PIANO_CANDIDATES; }
    // slot of a candidate = how many candidates precede it in (bucket, insertion index) order
    for (int a = tid; a < c; a += blockDim.x) {
        const sfb_piano_note n = notes[index[a]];
        int slot = 0;
        for (int b = 0; b < c; b++) {
            const int ob = notes[index[b]].order;
            slot += (bucket[b] < bucket[a] || (bucket[b] == bucket[a] && ob < n.order)) ? 1 : 0;
        }
        if (slot < PIANO_MAX_ROLLING)
            roll[size_t(p)*PIANO_MAX_ROLLING + slot] = make_float4(float(n.start), float(n.end), float(n.channel), float(n.velocity));
    }
}

extern "C" int sfb_piano_track(sfb_ctx* ctx, const sfb_piano_note* notes_dev, const int32_t* pitch_offset_dev,
                               const double* time_dev, int n_frames, double time_offset, double roll_time,
                               double lookup_time, double release_before_end, int global_min, int global_max,
                               float* key_target_out_dev, float* channel_out_dev, int32_t* upcoming_out_dev) {
    SFB_REQUIRE(ctx && pitch_offset_dev && time_dev && key_target_out_dev && channel_out_dev && upcoming_out_dev,
        "sfb_piano_track: null argument");
    SFB_REQUIRE(n_frames > 0, "sfb_piano_track: no frames");
    piano_track_kernel<<<n_frames, PIANO_MAX_NOTE, 0, ctx->stream>>>(notes_dev, pitch_offset_dev, time_dev, n_frames,
        time_offset, roll_time, lookup_time, release_before_end, global_min, global_max,
        key_target_out_dev, channel_out_dev, upcoming_out_dev);
    SFB_LAUNCH_CHECK(ctx);
    return SFB_OK;
}

extern "C" int sfb_piano_roll(sfb_ctx* ctx, const sfb_piano_note* notes_dev, const int32_t* pitch_offset_dev,
                              double time, double roll_time, double lookup_time, int global_min, int global_max,
                              float* roll_out_dev, int32_t* overflow_dev) {
    SFB_REQUIRE(ctx && pitch_offset_dev && roll_out_dev && overflow_dev, "sfb_piano_roll: null argument");
    piano_roll_kernel<<<PIANO_MAX_NOTE, 256, 0, ctx->stream>>>(notes_dev, pitch_offset_dev, time, roll_time, lookup_time,
        global_min, global_max, reinterpret_cast<float4*>(roll_out_dev), overflow_dev);
    SFB_LAUNCH_CHECK(ctx);
    return SFB_OK;
}
