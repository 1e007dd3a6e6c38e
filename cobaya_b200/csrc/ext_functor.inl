// ext_functor.inl -- run-time compilation of external likelihood functions (included by
// engine.cu).  NVRTC is loaded with dlopen at the first use, so that the library itself links
// nothing but the CUDA runtime; the compiled code is loaded through the runtime's library API
// (cudaLibraryLoadData / cudaLibraryGetKernel) and launched with cudaLaunchKernel.
#include <dlfcn.h>

#include <map>
#include <mutex>

namespace {

typedef struct _nvrtcProgram *nvrtcProgram_t;
struct NvrtcApi {
    void *lib = nullptr;
    int (*CreateProgram)(nvrtcProgram_t *, const char *, const char *, int, const char *const *,
                         const char *const *) = nullptr;
    int (*CompileProgram)(nvrtcProgram_t, int, const char *const *) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram_t, size_t *) = nullptr;
    int (*GetProgramLog)(nvrtcProgram_t, char *) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram_t, size_t *) = nullptr;
    int (*GetCUBIN)(nvrtcProgram_t, char *) = nullptr;
    int (*GetPTXSize)(nvrtcProgram_t, size_t *) = nullptr;
    int (*GetPTX)(nvrtcProgram_t, char *) = nullptr;
    int (*DestroyProgram)(nvrtcProgram_t *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string error;
};

NvrtcApi &nvrtc_api() {
    static NvrtcApi api;
    if (api.lib || !api.error.empty()) return api;
    const char *names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"};
    if (const char *env = getenv("CB2_NVRTC_LIB")) api.lib = dlopen(env, RTLD_NOW | RTLD_LOCAL);
    for (const char *nm : names) {
        if (api.lib) break;
        api.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
    }
    if (!api.lib) {
        api.error = "NVRTC not found (libnvrtc.so.12; set CB2_NVRTC_LIB)";
        return api;
    }
#define CB2_SYM(field, name)                                                   \
    *(void **)(&api.field) = dlsym(api.lib, name);                             \
    if (!api.field) api.error = std::string("NVRTC symbol missing: ") + name;
    CB2_SYM(CreateProgram, "nvrtcCreateProgram")
    CB2_SYM(CompileProgram, "nvrtcCompileProgram")
    CB2_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    CB2_SYM(GetProgramLog, "nvrtcGetProgramLog")
    CB2_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    CB2_SYM(GetCUBIN, "nvrtcGetCUBIN")
    CB2_SYM(GetPTXSize, "nvrtcGetPTXSize")
    CB2_SYM(GetPTX, "nvrtcGetPTX")
    CB2_SYM(DestroyProgram, "nvrtcDestroyProgram")
    CB2_SYM(GetErrorString, "nvrtcGetErrorString")
#undef CB2_SYM
    if (!api.error.empty()) { dlclose(api.lib); api.lib = nullptr; }
    return api;
}

bool valid_identifier(const char *s) {
    if (!s || !*s) return false;
    if (!(isalpha((unsigned char)*s) || *s == '_')) return false;
    for (const char *p = s; *p; ++p)
        if (!(isalnum((unsigned char)*p) || *p == '_')) return false;
    return true;
}

// the translation unit handed to NVRTC: the user's source followed by the evaluation kernel
std::string ext_translation_unit(const char *source, const char *fn, int d,
                                 const int32_t *idx) {
    std::string s;
    s += "#define CB2_EXT_DIM " + std::to_string(d) + "\n";
    s += "#line 1 \"external_likelihood.cu\"\n";
    s += source;
    s += "\n#line 1 \"cb2_ext_eval.cu\"\n";
    s += "extern \"C\" __device__ double ";
    s += fn;
    s += "(const double *p, int n);\n";
    s += "extern \"C\" __global__ void cb2_ext_eval(const double *__restrict__ X, long long n,\n"
         "        int D, double *__restrict__ out, int stride, int off) {\n"
         "    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;\n"
         "    if (c >= n) return;\n"
         "    const int idx[CB2_EXT_DIM] = {";
    for (int i = 0; i < d; ++i) s += (i ? ", " : "") + std::to_string(idx[i]);
    s += "};\n"
         "    double p[CB2_EXT_DIM];\n"
         "#pragma unroll\n"
         "    for (int i = 0; i < CB2_EXT_DIM; ++i) p[i] = X[c * D + idx[i]];\n"
         "    out[c * stride + off] = ";
    s += fn;
    s += "(p, CB2_EXT_DIM);\n}\n";
    return s;
}

// source -> device code (CUBIN for sm_100a, PTX if the CUBIN is not available); no GPU needed
int ext_compile(const char *source, const char *fn, int d, const int32_t *idx,
                std::vector<char> &image, std::string &log) {
    NvrtcApi &api = nvrtc_api();
    if (!api.lib) { log = api.error; return -6; }
    const std::string tu = ext_translation_unit(source, fn, d, idx);
    nvrtcProgram_t prog = nullptr;
    int rc = api.CreateProgram(&prog, tu.c_str(), "cb2_external.cu", 0, nullptr, nullptr);
    if (rc) { log = std::string("nvrtcCreateProgram: ") + api.GetErrorString(rc); return -6; }
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=true",
                          "-default-device"};
    rc = api.CompileProgram(prog, 4, opts);
    size_t ls = 0;
    api.GetProgramLogSize(prog, &ls);
    if (ls > 1) {
        std::vector<char> buf(ls + 1, 0);
        api.GetProgramLog(prog, buf.data());
        log = buf.data();
    }
    if (rc) {
        if (log.empty()) log = api.GetErrorString(rc);
        api.DestroyProgram(&prog);
        return -7;
    }
    size_t n = 0;
    if (api.GetCUBINSize(prog, &n) == 0 && n > 0) {
        image.resize(n);
        api.GetCUBIN(prog, image.data());
    } else if (api.GetPTXSize(prog, &n) == 0 && n > 0) {
        image.resize(n);
        api.GetPTX(prog, image.data());
    } else {
        log += " (no device code produced)";
        api.DestroyProgram(&prog);
        return -7;
    }
    api.DestroyProgram(&prog);
    return 0;
}

// ---- the general step kernel with the user's functions inlined ----------------------------
#include "embedded_sources.inc"

struct FusedUser {
    std::string source, name;
    int dim;
    int user_id;   // index the kernel dispatches on (LikeDev.n_modes of a kind-3 likelihood)
};

// k_step_general (csrc/kernels_general.cuh, compiled again from the copy embedded in the
// library) + the user's likelihood functions in ONE translation unit: the functions are
// inlined where warp_logpost evaluates a kind-3 likelihood, so a whole window -- Metropolis or
// dragging -- is one launch, as for the built-in likelihoods.
std::string fused_translation_unit(const std::vector<FusedUser> &users) {
    std::string s;
    s += "#define CB2_NVRTC_USER 1\n#include \"kernels_general.cuh\"\n";
    for (const FusedUser &u : users) {
        s += "#undef CB2_EXT_DIM\n#define CB2_EXT_DIM " + std::to_string(u.dim) + "\n";
        s += "namespace cb2_user_" + std::to_string(u.user_id) + " {\n";
        s += "#line 1 \"external_" + u.name + ".cu\"\n";
        s += u.source;
        s += "\n}\n";
    }
    s += "#line 1 \"cb2_user_dispatch.cu\"\n";
    s += "__device__ double cb2_user_like(int user_id, const double *p, int n) {\n"
         "    switch (user_id) {\n";
    for (const FusedUser &u : users)
        s += "        case " + std::to_string(u.user_id) + ": return cb2_user_" +
             std::to_string(u.user_id) + "::" + u.name + "(p, n);\n";
    s += "    }\n    return 0.0;\n}\n";
    s += "extern \"C\" __global__ void __launch_bounds__(256) cb2_step_user(ModelDev M, "
         "ChainState S, WindowDev W,\n        StepSmem L, long long n_chains, "
         "unsigned long long t0, int n_steps) {\n"
         "    step_general_body(M, S, W, L, n_chains, t0, n_steps);\n}\n";
    return s;
}

int fused_compile(const std::vector<FusedUser> &users, std::vector<char> &image,
                  std::string &log) {
    NvrtcApi &api = nvrtc_api();
    if (!api.lib) { log = api.error; return -6; }
    const std::string tu = fused_translation_unit(users);
    // engines of one process that share their functions share the compiled image (the compile
    // takes ~6 s: the whole general step kernel)
    static std::map<std::string, std::vector<char>> cache;
    static std::mutex cache_mutex;
    {
        std::lock_guard<std::mutex> g(cache_mutex);
        auto it = cache.find(tu);
        if (it != cache.end()) { image = it->second; return 0; }
    }
    const int nh = (int)(sizeof(cb2_embedded_names) / sizeof(cb2_embedded_names[0]));
    nvrtcProgram_t prog = nullptr;
    int rc = api.CreateProgram(&prog, tu.c_str(), "cb2_step_user.cu", nh, cb2_embedded_sources,
                               cb2_embedded_names);
    if (rc) { log = std::string("nvrtcCreateProgram: ") + api.GetErrorString(rc); return -6; }
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=true",
                          "-default-device", "-lineinfo"};
    rc = api.CompileProgram(prog, 5, opts);
    size_t ls = 0;
    api.GetProgramLogSize(prog, &ls);
    if (ls > 1) {
        std::vector<char> buf(ls + 1, 0);
        api.GetProgramLog(prog, buf.data());
        log = buf.data();
    }
    if (rc) {
        if (log.empty()) log = api.GetErrorString(rc);
        api.DestroyProgram(&prog);
        return -7;
    }
    size_t n = 0;
    if (api.GetCUBINSize(prog, &n) == 0 && n > 0) {
        image.resize(n);
        api.GetCUBIN(prog, image.data());
    } else if (api.GetPTXSize(prog, &n) == 0 && n > 0) {
        image.resize(n);
        api.GetPTX(prog, image.data());
    } else {
        log += " (no device code produced)";
        api.DestroyProgram(&prog);
        return -7;
    }
    api.DestroyProgram(&prog);
    {
        std::lock_guard<std::mutex> g(cache_mutex);
        cache[tu] = image;
    }
    return 0;
}

}  // namespace
