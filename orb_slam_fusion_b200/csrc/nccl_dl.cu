// nccl_dl.cu -- run-time binding of NCCL (see nccl_dl.cuh).
#include "nccl_dl.cuh"

#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>

namespace orbx {

namespace {
NcclApi g_api;
std::once_flag g_once;
bool g_ok = false;

void bind() {
  const char* forced = getenv("ORBX_NCCL_LIB");
  void* lib = nullptr;
  if (forced && forced[0]) lib = dlopen(forced, RTLD_NOW | RTLD_LOCAL);
  // the NCCL this process already uses (torch's bundled copy), else the system one
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!lib) {
    snprintf(g_api.err, sizeof(g_api.err), "cannot load libnccl.so.2: %s", dlerror());
    return;
  }
  g_api.lib = lib;
  struct { const char* name; void** slot; } syms[] = {
      {"ncclGetErrorString", (void**)&g_api.GetErrorString}, {"ncclGetUniqueId", (void**)&g_api.GetUniqueId},
      {"ncclCommInitRank", (void**)&g_api.CommInitRank},     {"ncclCommDestroy", (void**)&g_api.CommDestroy},
      {"ncclCommCount", (void**)&g_api.CommCount},           {"ncclCommUserRank", (void**)&g_api.CommUserRank},
      {"ncclAllGather", (void**)&g_api.AllGather},           {"ncclGetVersion", (void**)&g_api.GetVersion},
  };
  for (auto& s : syms) {
    *s.slot = dlsym(lib, s.name);
    if (!*s.slot) {
      snprintf(g_api.err, sizeof(g_api.err), "libnccl lacks %s", s.name);
      return;
    }
  }
  g_ok = true;
}
}  // namespace

const NcclApi* nccl_api(const char** why) {
  std::call_once(g_once, bind);
  if (!g_ok && why) *why = g_api.err;
  return g_ok ? &g_api : nullptr;
}

}  // namespace orbx
