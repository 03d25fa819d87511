/* sw_jit.cu -- see sw_jit.h. */
#include "sw_jit.h"

#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {

const unsigned char kStripSource[] = {
#include "sw_strip_src.inc"
};

// NVRTC without its header: only what is used here
typedef struct _nvrtcProgram *nvrtcProgram;
typedef int nvrtcResult;
struct Nvrtc {
    void *so = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram *) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
    nvrtcResult (*AddNameExpression)(nvrtcProgram, const char *) = nullptr;
    nvrtcResult (*GetLoweredName)(nvrtcProgram, const char *, const char **) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
    bool ok = false;
};

Nvrtc &nvrtc()
{
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
        for (const char *nm : names) {
            n.so = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (n.so) break;
        }
        if (!n.so) return;
#define SW_SYM(f) *(void **)(&n.f) = dlsym(n.so, "nvrtc" #f)
        SW_SYM(CreateProgram); SW_SYM(DestroyProgram); SW_SYM(CompileProgram); SW_SYM(GetCUBINSize); SW_SYM(GetCUBIN);
        SW_SYM(AddNameExpression); SW_SYM(GetLoweredName); SW_SYM(GetProgramLogSize); SW_SYM(GetProgramLog);
#undef SW_SYM
        n.ok = n.CreateProgram && n.DestroyProgram && n.CompileProgram && n.GetCUBINSize && n.GetCUBIN &&
               n.AddNameExpression && n.GetLoweredName && n.GetProgramLogSize && n.GetProgramLog;
    });
    return n;
}

uint64_t fnv1a(const void *p, size_t n, uint64_t h = 1469598103934665603ull)
{
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

std::string cache_dir()
{
    std::string d;
    if (const char *e = std::getenv("SW_B200_JIT_CACHE")) d = e;
    else if (const char *h = std::getenv("HOME")) d = std::string(h) + "/.cache/sw_b200";
    else d = "/tmp/sw_b200_cache";
    return d;
}

void mkdirs(const std::string &d)
{
    std::string cur;
    for (size_t i = 0; i <= d.size(); ++i) {
        if (i == d.size() || d[i] == '/') { if (!cur.empty()) mkdir(cur.c_str(), 0755); }
        if (i < d.size()) cur.push_back(d[i]);
    }
}

struct Entry { void *kernel = nullptr; std::string why; };
std::map<std::string, Entry> g_cache;
std::mutex g_mu;

void set_msg(char *msg, size_t cap, const std::string &s)
{
    if (msg && cap) { std::snprintf(msg, cap, "%s", s.c_str()); }
}

}  // namespace

int sw_jit_available(void) { return nvrtc().ok ? 1 : 0; }

void *sw_jit_strip_kernel(const SwStripVariant *v, int goe, int ge, char *msg, size_t msg_cap)
{
    if (!v) return nullptr;
    const int rs = v->R / v->S;
    char expr[256];
    std::snprintf(expr, sizeof expr, "swk::sw_strip_kernel<%d, %d, %d, swk::ArithS16, false, %d, %d, %d, %d, false, %d, %d>", rs, v->S, v->G,
                  v->block_threads, v->min_blocks, goe, ge, v->U, v->FL);
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_cache.find(expr);
    if (it != g_cache.end()) { set_msg(msg, msg_cap, it->second.why); return it->second.kernel; }
    Entry &ent = g_cache[expr];

    // prelude: NVRTC has no <stdint.h>
    static const char prelude[] =
        "#define SW_JIT_BUILD 1\n"
        "typedef unsigned char uint8_t; typedef signed char int8_t; typedef unsigned short uint16_t; typedef short int16_t;\n"
        "typedef unsigned int uint32_t; typedef int int32_t; typedef unsigned long long uint64_t; typedef long long int64_t;\n";
    std::string src = std::string(prelude) + (const char *)kStripSource;
    uint64_t hsh = fnv1a(src.data(), src.size());
    hsh = fnv1a(expr, std::strlen(expr), hsh);
#ifdef SW_BOUNDS_CHECK
    hsh = fnv1a("check", 5, hsh);
#endif
    char fname[64];
    std::snprintf(fname, sizeof fname, "/strip_%016llx.cubin", (unsigned long long)hsh);
    const std::string dir = cache_dir(), path = dir + fname, lpath = path + ".name";

    std::vector<char> cubin;
    std::string lowered;
    if (FILE *f = std::fopen(path.c_str(), "rb")) {
        std::fseek(f, 0, SEEK_END);
        long sz = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        if (sz > 0) { cubin.resize((size_t)sz); if (std::fread(cubin.data(), 1, (size_t)sz, f) != (size_t)sz) cubin.clear(); }
        std::fclose(f);
        if (FILE *g = std::fopen(lpath.c_str(), "r")) {
            char buf[1024];
            if (std::fgets(buf, sizeof buf, g)) { lowered = buf; while (!lowered.empty() && (lowered.back() == '\n')) lowered.pop_back(); }
            std::fclose(g);
        }
        if (lowered.empty()) cubin.clear();
    }
    if (cubin.empty()) {
        Nvrtc &n = nvrtc();
        if (!n.ok) { ent.why = "NVRTC (libnvrtc.so.12) could not be loaded"; set_msg(msg, msg_cap, ent.why); return nullptr; }
        nvrtcProgram prog = nullptr;
        if (n.CreateProgram(&prog, src.c_str(), "sw_strip_jit.cu", 0, nullptr, nullptr) != 0) {
            ent.why = "nvrtcCreateProgram failed"; set_msg(msg, msg_cap, ent.why); return nullptr;
        }
        n.AddNameExpression(prog, expr);
        // -default-device: NVRTC takes the kernel's generic lambda (no execution-space annotation, as
        // nvcc wants it inside device code) for a host function otherwise
        const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-DSW_JIT_BUILD=1", "-default-device",
#ifdef SW_BOUNDS_CHECK
                              "-DSW_BOUNDS_CHECK=1",
#endif
        };
        const nvrtcResult rc = n.CompileProgram(prog, (int)(sizeof(opts) / sizeof(opts[0])), opts);
        if (rc != 0) {
            size_t ls = 0;
            n.GetProgramLogSize(prog, &ls);
            std::string log(ls, '\0');
            if (ls) n.GetProgramLog(prog, &log[0]);
            ent.why = "NVRTC compile failed: " + log.substr(0, 400);
            n.DestroyProgram(&prog);
            set_msg(msg, msg_cap, ent.why);
            return nullptr;
        }
        const char *low = nullptr;
        size_t cs = 0;
        if (n.GetLoweredName(prog, expr, &low) != 0 || !low || n.GetCUBINSize(prog, &cs) != 0 || cs == 0) {
            ent.why = "NVRTC produced no cubin / lowered name";
            n.DestroyProgram(&prog);
            set_msg(msg, msg_cap, ent.why);
            return nullptr;
        }
        lowered = low;
        cubin.resize(cs);
        n.GetCUBIN(prog, cubin.data());
        n.DestroyProgram(&prog);
        // best-effort disk cache (atomic rename)
        mkdirs(dir);
        const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
        if (FILE *f = std::fopen(tmp.c_str(), "wb")) {
            const bool okw = std::fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
            std::fclose(f);
            if (okw) {
                if (FILE *g = std::fopen(lpath.c_str(), "w")) { std::fprintf(g, "%s\n", lowered.c_str()); std::fclose(g); }
                std::rename(tmp.c_str(), path.c_str());
            } else {
                std::remove(tmp.c_str());
            }
        }
    }
    cudaLibrary_t lib = nullptr;
    cudaError_t e = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    cudaKernel_t k = nullptr;
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&k, lib, lowered.c_str());
    if (e != cudaSuccess) {
        cudaGetLastError();
        ent.why = std::string("loading the specialised cubin failed: ") + cudaGetErrorString(e);
        set_msg(msg, msg_cap, ent.why);
        return nullptr;
    }
    ent.kernel = (void *)k;
    ent.why = "ok";
    set_msg(msg, msg_cap, ent.why);
    return ent.kernel;
}
