// jit.cu — run-time compiled fragment programs: the `compile()` half of ShaderProgram for fragment text this backend
// has no ahead-of-time kernel for (reference: shaderflow/shader.py:313-349 hands any GLSL to the GL driver; here the
// Python side translates it to CUDA — shaderflow_b200/glsl — and this unit turns that source into kernels).
//   sfb_jit_compile    NVRTC → SASS for sm_100a. Needs no GPU (the CPU test-suite compiles through it); libnvrtc is
//                      loaded on first use, so the library itself does not depend on it.
//   sfb_program_load   cudaLibraryLoadData + the two entry kernels (jit/jit_kernels.cuh) → a scene id ≥
//                      SFB_SCENE_PROGRAM_BASE that sfb_render_screen / _target / _frame accept like a built-in scene.
#include "sfb_internal.h"

#include <dlfcn.h>
#include <mutex>
#include <nvrtc.h>
#include <stdlib.h>

namespace {

struct Nvrtc {
    void* handle = nullptr;
    nvrtcResult (*create)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*destroy)(nvrtcProgram*) = nullptr;
    nvrtcResult (*compile)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*log_size)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*log)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*cubin_size)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*cubin)(nvrtcProgram, char*) = nullptr;
    const char* (*error_string)(nvrtcResult) = nullptr;
};

const Nvrtc* nvrtc() {
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
            n.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (n.handle) break;
        }
        if (!n.handle) return;
        #define SFB_SYM(field, symbol) n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.handle, symbol))
        SFB_SYM(create, "nvrtcCreateProgram"); SFB_SYM(destroy, "nvrtcDestroyProgram"); SFB_SYM(compile, "nvrtcCompileProgram");
        SFB_SYM(log_size, "nvrtcGetProgramLogSize"); SFB_SYM(log, "nvrtcGetProgramLog");
        SFB_SYM(cubin_size, "nvrtcGetCUBINSize"); SFB_SYM(cubin, "nvrtcGetCUBIN"); SFB_SYM(error_string, "nvrtcGetErrorString");
        #undef SFB_SYM
        if (!(n.create && n.destroy && n.compile && n.log_size && n.log && n.cubin_size && n.cubin && n.error_string)) {
            dlclose(n.handle); n.handle = nullptr;
        }
    });
    return n.handle ? &n : nullptr;
}

struct Program {
    bool live = false;
    int device = 0, n_samplers = 0;
    cudaLibrary_t library = nullptr;
    cudaKernel_t screen = nullptr, frame = nullptr;
};
std::mutex g_mutex;
std::vector<Program> g_programs;

}  // namespace

extern "C" int sfb_jit_compile(const char* source, const char* const* header_names, const char* const* header_sources,
                               int n_headers, int flags, void** image, size_t* image_bytes, char** log) {
    SFB_REQUIRE(source && image && image_bytes, "sfb_jit_compile: null argument");
    SFB_REQUIRE(n_headers >= 0 && (n_headers == 0 || (header_names && header_sources)), "sfb_jit_compile: bad header table");
    *image = nullptr; *image_bytes = 0;
    if (log) *log = nullptr;
    const Nvrtc* rt = nvrtc();
    if (!rt) SFB_FAIL(SFB_ENOTFOUND, "sfb_jit_compile: libnvrtc.so.12 could not be loaded (%s)", dlerror());
    nvrtcProgram program = nullptr;
    nvrtcResult rc = rt->create(&program, source, "fragment.cu", n_headers, header_sources, header_names);
    if (rc != NVRTC_SUCCESS) SFB_FAIL(SFB_ECUDA, "nvrtcCreateProgram: %s", rt->error_string(rc));
    const char* options[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "-default-device",     // sfb200.h declares the (host) C ABI
                             (flags & SFB_JIT_FMAD) ? "--fmad=true" : "--fmad=false"};
    rc = rt->compile(program, int(sizeof(options)/sizeof(options[0])), options);
    size_t log_bytes = 0;
    if (log && rt->log_size(program, &log_bytes) == NVRTC_SUCCESS && log_bytes > 1) {
        *log = static_cast<char*>(malloc(log_bytes));
        if (*log && rt->log(program, *log) != NVRTC_SUCCESS) { free(*log); *log = nullptr; }
    }
    if (rc != NVRTC_SUCCESS) {
        const char* why = rt->error_string(rc);
        rt->destroy(&program);
        SFB_FAIL(SFB_EINVAL, "sfb_jit_compile: %s (see the compile log)", why);
    }
    size_t bytes = 0;
    rc = rt->cubin_size(program, &bytes);
    void* out = (rc == NVRTC_SUCCESS && bytes) ? malloc(bytes) : nullptr;
    if (out) rc = rt->cubin(program, static_cast<char*>(out));
    rt->destroy(&program);
    if (!out || rc != NVRTC_SUCCESS) { free(out); SFB_FAIL(SFB_ECUDA, "sfb_jit_compile: no SASS image (%s)", rt->error_string(rc)); }
    *image = out; *image_bytes = bytes;
    return SFB_OK;
}

extern "C" void sfb_jit_free(void* p) { free(p); }

extern "C" int sfb_program_load(sfb_ctx* ctx, const void* image, size_t image_bytes, int n_samplers, int* scene) {
    SFB_REQUIRE(ctx && image && image_bytes && scene, "sfb_program_load: null argument");
    SFB_REQUIRE(n_samplers >= 0 && n_samplers <= SFB_MAX_SAMPLERS, "sfb_program_load: %d samplers", n_samplers);
    Program p;
    p.device = ctx->device; p.n_samplers = n_samplers;
    SFB_CUDA(cudaLibraryLoadData(&p.library, image, nullptr, nullptr, 0, nullptr, nullptr, 0));
    cudaError_t e = cudaLibraryGetKernel(&p.screen, p.library, "sfb_jit_screen");
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&p.frame, p.library, "sfb_jit_frame");
    if (e != cudaSuccess) {
        cudaLibraryUnload(p.library);
        SFB_FAIL(SFB_ECUDA, "sfb_program_load: the image has no sfb_jit_screen / sfb_jit_frame kernels (%s)", cudaGetErrorString(e));
    }
    p.live = true;
    std::lock_guard<std::mutex> lock(g_mutex);
    size_t slot = 0;
    while (slot < g_programs.size() && g_programs[slot].live) slot++;
    if (slot == g_programs.size()) g_programs.emplace_back();
    g_programs[slot] = p;
    *scene = SFB_SCENE_PROGRAM_BASE + int(slot);
    return SFB_OK;
}

extern "C" int sfb_program_unload(sfb_ctx* ctx, int scene) {
    SFB_REQUIRE(ctx, "sfb_program_unload: null ctx");
    std::lock_guard<std::mutex> lock(g_mutex);
    const int slot = scene - SFB_SCENE_PROGRAM_BASE;
    SFB_REQUIRE(slot >= 0 && slot < int(g_programs.size()) && g_programs[slot].live, "sfb_program_unload: bad program %d", scene);
    cudaStreamSynchronize(ctx->stream);
    cudaLibraryUnload(g_programs[slot].library);
    g_programs[slot] = Program();
    return SFB_OK;
}

// samplers the program reads, or -1 when `scene` is not a loaded program of this device
int sfb_program_samplers(int scene, int device) {
    std::lock_guard<std::mutex> lock(g_mutex);
    const int slot = scene - SFB_SCENE_PROGRAM_BASE;
    if (slot < 0 || slot >= int(g_programs.size()) || !g_programs[slot].live || g_programs[slot].device != device) return -1;
    return g_programs[slot].n_samplers;
}

// kind 0: one thread per fragment of the Wr x Hr target; kind 1: fused ssaa + final pass per output pixel
int sfb_program_launch(int scene, int kind, const RenderParams& P, cudaStream_t stream) {
    cudaKernel_t kernel = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        const int slot = scene - SFB_SCENE_PROGRAM_BASE;
        SFB_REQUIRE(slot >= 0 && slot < int(g_programs.size()) && g_programs[slot].live, "bad program %d", scene);
        kernel = kind ? g_programs[slot].frame : g_programs[slot].screen;
    }
    const int w = kind ? P.W : P.Wr, h = kind ? P.H : P.Hr;
    dim3 block(32, 8), grid((w + 31)/32, (h + 7)/8);
    void* args[] = {const_cast<RenderParams*>(&P)};
    SFB_CUDA(cudaLaunchKernel(reinterpret_cast<const void*>(kernel), grid, block, args, 0, stream));
    return SFB_OK;
}
