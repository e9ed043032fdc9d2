// napi_harness.cc -- TEST INFRASTRUCTURE: drives js/src/binding.cc through the mock N-API
// (napi_mock.cc) in the order js/lib/gpuSnapshotStage.js does:
//   open({...}) -> watch(h, wake) -> [write(h, Buffer) until 0 / all] ... flush(h)
//   on every wake: peek/consume until null; 'eof' -> stats, endChecksum, unwatch, close.
// usage: napi_harness <mode 0..4> <input file> <output file> [chunk bytes]
// prints one JSON line; exit 0 ok, 3 = an exported function threw (code/message printed).
#include "napi_mock.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static napi_env env;
static napi_value addon;

static bool threw(const char *where)
{
	std::string code, msg;
	if (!mock_exception(env, &code, &msg)) return false;
	printf("{\"threw\": \"%s\", \"code\": \"%s\", \"message\": \"%s\"}\n", where, code.c_str(), msg.c_str());
	return true;
}

static napi_value call(const char *name, std::vector<napi_value> args)
{
	napi_value fn = mock_get(addon, name);
	if (fn == NULL) { fprintf(stderr, "addon has no export %s\n", name); exit(2); }
	return mock_call(env, fn, args);
}

int main(int argc, char **argv)
{
	if (argc < 4) { fprintf(stderr, "usage: %s mode in out [chunk]\n", argv[0]); return 2; }
	const int mode = atoi(argv[1]);
	const size_t chunk = argc > 4 ? (size_t)atol(argv[4]) : (1u << 20);
	FILE *fi = fopen(argv[2], "rb"), *fo = fopen(argv[3], "wb");
	if (!fi || !fo) { perror("open"); return 2; }
	fseek(fi, 0, SEEK_END);
	const size_t total = (size_t)ftell(fi);
	fseek(fi, 0, SEEK_SET);
	std::vector<uint8_t> in(total);
	if (total && fread(in.data(), 1, total, fi) != total) { perror("read"); return 2; }

	env = mock_env_new();
	addon = mock_load_addon(env);
	napi_value cfg = mock_object(env);
	mock_set(cfg, "mode", mock_number(env, mode));
	mock_set(cfg, "device", mock_number(env, 0));
	mock_set(cfg, "ringBytes", mock_number(env, 8 << 20));
	mock_set(cfg, "outRingBytes", mock_number(env, 8 << 20));
	mock_set(cfg, "batchBytes", mock_number(env, 1 << 20));
	napi_value h = call("open", { cfg });
	if (threw("open")) return 3;

	int wakes = 0;
	napi_value watcher = call("watch", { h, mock_function(env, [&] { wakes++; }) });
	if (threw("watch")) return 3;

	size_t fed = 0, out_bytes = 0, zero_writes = 0;
	bool flushed = false, eof = false;
	int idle = 0;
	while (!eof) {
		// Transform._write: offer the next chunk; 0 accepted == ring full, retry on wake-up
		while (fed < total) {
			const size_t n = std::min(chunk, total - fed);
			napi_value r = call("write", { h, mock_buffer(env, in.data() + fed, n) });
			if (threw("write")) { call("unwatch", { watcher }); call("close", { h }); return 3; }
			const size_t acc = (size_t)r->num;
			fed += acc;
			if (acc < n) { if (acc == 0) zero_writes++; break; }
		}
		if (fed == total && !flushed) {               // Transform._flush
			call("flush", { h });
			if (threw("flush")) { call("unwatch", { watcher }); call("close", { h }); return 3; }
			flushed = true;
		}
		// _drain: runs when the watcher's function fires on the "event loop"
		const int fired = mock_run_loop(env, 200);
		idle = fired ? 0 : idle + 1;
		if (idle > 300) { printf("{\"error\": \"no wake-up for 60 s\"}\n"); return 4; }
		for (;;) {
			napi_value ab = call("peek", { h });
			if (threw("peek")) { call("unwatch", { watcher }); call("close", { h }); return 3; }
			if (ab->kind == MK_NULL) break;
			if (ab->kind == MK_STRING && ab->str == "eof") { eof = true; break; }
			if (ab->kind != MK_ARRAYBUFFER) { fprintf(stderr, "peek returned kind %d\n", ab->kind); return 2; }
			fwrite(ab->ptr, 1, ab->len, fo);       // Buffer.from(Buffer.from(ab)): copy out, then release
			out_bytes += ab->len;
			call("consume", { h, mock_number(env, (double)ab->len) });
			if (threw("consume")) { call("unwatch", { watcher }); call("close", { h }); return 3; }
		}
	}
	fclose(fo);
	napi_value st = call("stats", { h });
	if (threw("stats")) return 3;
	napi_value ck = call("endChecksum", { h });
	if (threw("endChecksum")) return 3;
	call("unwatch", { watcher });
	call("close", { h });
	printf("{\"ok\": true, \"fed\": %zu, \"out\": %zu, \"wakes\": %d, \"ring_full\": %zu, "
	    "\"bytesIn\": %.0f, \"bytesOut\": %.0f, \"records\": %.0f, \"lz4Encoded\": %.0f, "
	    "\"endChecksum\": [\"%016llx\", \"%016llx\", \"%016llx\", \"%016llx\"]}\n",
	    fed, out_bytes, wakes, zero_writes, mock_get(st, "bytesIn")->num, mock_get(st, "bytesOut")->num,
	    mock_get(st, "records")->num, mock_get(st, "lz4Encoded")->num,
	    (unsigned long long)ck->elems[0]->big, (unsigned long long)ck->elems[1]->big,
	    (unsigned long long)ck->elems[2]->big, (unsigned long long)ck->elems[3]->big);
	return 0;
}
