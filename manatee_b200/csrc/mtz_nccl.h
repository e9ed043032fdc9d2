// mtz_nccl.h -- NCCL, bound at run time.
//
// The library owns its NCCL communicators (fan-out broadcast over the device group, the shard
// exchange between ranks) but does not carry a link-time dependency on one particular libnccl:
// a host process may already have one loaded under the same SONAME (PyTorch bundles its own
// libnccl.so.2, newer than the system's, and would fail to import behind an older copy), and a
// process that never fans out needs none.  The few entry points used are resolved with dlopen on
// first use: $MTZ_NCCL_LIB if set, else whatever libnccl.so.2 the process already holds, else the
// system's.  Their signatures are unchanged across NCCL 2.x; the types come from <nccl.h>.
#pragma once
#include <nccl.h>
#ifndef MTZ_HOST_EMUL          // tests/emul supplies an in-process stub nccl.h instead
#include <dlfcn.h>
#include <stdlib.h>
#include <string>

namespace mtz {

struct NcclApi {
	decltype(&::ncclGetErrorString) GetErrorString = nullptr;
	decltype(&::ncclGetUniqueId) GetUniqueId = nullptr;
	decltype(&::ncclCommInitAll) CommInitAll = nullptr;
	decltype(&::ncclCommInitRank) CommInitRank = nullptr;
	decltype(&::ncclCommDestroy) CommDestroy = nullptr;
	decltype(&::ncclGroupStart) GroupStart = nullptr;
	decltype(&::ncclGroupEnd) GroupEnd = nullptr;
	decltype(&::ncclBroadcast) Broadcast = nullptr;
	decltype(&::ncclAllGather) AllGather = nullptr;
	decltype(&::ncclSend) Send = nullptr;
	decltype(&::ncclRecv) Recv = nullptr;
	decltype(&::ncclGetVersion) GetVersion = nullptr;
	bool ok = false;
	std::string err, path;
	NcclApi()
	{
		void *h = nullptr;
		const char *env = getenv("MTZ_NCCL_LIB");
		if (env && *env) { h = dlopen(env, RTLD_NOW | RTLD_LOCAL); path = env; }
		if (h == nullptr) { h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD); path = "libnccl.so.2 (already loaded)"; }
		if (h == nullptr) { h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL); path = "libnccl.so.2"; }
		if (h == nullptr) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
#define MTZ_SYM(field, name)                                                    \
		field = reinterpret_cast<decltype(field)>(dlsym(h, name));             \
		if (field == nullptr) { err = std::string("libnccl lacks ") + name; return; }
		MTZ_SYM(GetErrorString, "ncclGetErrorString")
		MTZ_SYM(GetUniqueId, "ncclGetUniqueId")
		MTZ_SYM(CommInitAll, "ncclCommInitAll")
		MTZ_SYM(CommInitRank, "ncclCommInitRank")
		MTZ_SYM(CommDestroy, "ncclCommDestroy")
		MTZ_SYM(GroupStart, "ncclGroupStart")
		MTZ_SYM(GroupEnd, "ncclGroupEnd")
		MTZ_SYM(Broadcast, "ncclBroadcast")
		MTZ_SYM(AllGather, "ncclAllGather")
		MTZ_SYM(Send, "ncclSend")
		MTZ_SYM(Recv, "ncclRecv")
		MTZ_SYM(GetVersion, "ncclGetVersion")
#undef MTZ_SYM
		ok = true;
	}
};

inline NcclApi &nccl_api() { static NcclApi a; return a; }
inline bool nccl_available(std::string *why)
{
	NcclApi &a = nccl_api();
	if (!a.ok && why) *why = a.err;
	return a.ok;
}

} // namespace mtz

#define ncclGetErrorString (mtz::nccl_api().GetErrorString)
#define ncclGetUniqueId    (mtz::nccl_api().GetUniqueId)
#define ncclCommInitAll    (mtz::nccl_api().CommInitAll)
#define ncclCommInitRank   (mtz::nccl_api().CommInitRank)
#define ncclCommDestroy    (mtz::nccl_api().CommDestroy)
#define ncclGroupStart     (mtz::nccl_api().GroupStart)
#define ncclGroupEnd       (mtz::nccl_api().GroupEnd)
#define ncclBroadcast      (mtz::nccl_api().Broadcast)
#define ncclAllGather      (mtz::nccl_api().AllGather)
#define ncclSend           (mtz::nccl_api().Send)
#define ncclRecv           (mtz::nccl_api().Recv)

#else
namespace mtz { inline bool nccl_available(std::string *) { return true; } }
#endif
