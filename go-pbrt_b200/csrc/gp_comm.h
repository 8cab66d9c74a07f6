// gp_comm.h — NCCL binding of libgopbrt_cuda.so: the one collective of the path, the film reduce that follows the last
// wavefront when a frame's samples are split across GPUs (north star: "per-GPU films are summed with one NCCL reduce over
// NVLink"; the reference merges its workers' FilmTiles on one host, film.go:115-132).
//
// libnccl is bound at run time (dlopen), not at link time: a host process that already carries a NCCL (PyTorch bundles its
// own libnccl.so.2) must not get a second copy with the same SONAME mapped under it, and single-GPU hosts need none at all.
// GOPBRT_NCCL_LIB overrides the library path.  Only the stable C API of nccl.h is used (declared here, so the build does not
// depend on the header's location).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <mutex>
#include <string>

namespace gpcomm {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;  // NCCL_UNIQUE_ID_BYTES = 128
enum { kNcclSuccess = 0, kNcclSum = 0, kNcclFloat64 = 8 };

struct Api {
  void* handle = nullptr;
  int (*GetVersion)(int*) = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string error;
  int version = 0;
};

// the process-wide binding; nullptr (and `error`) when libnccl cannot be bound — callers fail, there is no substitute path
inline Api* api(std::string& error) {
  static Api a;
  static std::once_flag once;
  std::call_once(once, []() {
    const char* names[4] = {getenv("GOPBRT_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
    // a NCCL the process already mapped wins (RTLD_NOLOAD), then the loader's search path
    for (int pass = 0; pass < 2 && !a.handle; pass++)
      for (int k = 0; k < 3 && !a.handle; k++)
        if (names[k] && names[k][0]) a.handle = dlopen(names[k], RTLD_NOW | RTLD_LOCAL | (pass == 0 ? RTLD_NOLOAD : 0));
    if (!a.handle) { a.error = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "dlopen failed"); return; }
    bool ok = true;
    auto sym = [&](const char* n) { void* p = dlsym(a.handle, n); if (!p) { ok = false; a.error = std::string("libnccl: missing symbol ") + n; } return p; };
    a.GetVersion = (int (*)(int*))sym("ncclGetVersion");
    a.GetUniqueId = (int (*)(ncclUniqueId*))sym("ncclGetUniqueId");
    a.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))sym("ncclCommInitRank");
    a.CommInitAll = (int (*)(ncclComm_t*, int, const int*))sym("ncclCommInitAll");
    a.CommDestroy = (int (*)(ncclComm_t))sym("ncclCommDestroy");
    a.Reduce = (int (*)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t))sym("ncclReduce");
    a.GroupStart = (int (*)())sym("ncclGroupStart");
    a.GroupEnd = (int (*)())sym("ncclGroupEnd");
    a.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    if (!ok) { dlclose(a.handle); a.handle = nullptr; return; }
    a.GetVersion(&a.version);
  });
  if (!a.handle) { error = a.error; return nullptr; }
  return &a;
}

}  // namespace gpcomm
