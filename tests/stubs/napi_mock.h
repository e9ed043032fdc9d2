// napi_mock.h -- TEST INFRASTRUCTURE (see napi_mock.cc): the harness-side view of the mock
#pragma once
#include "node_api.h"
#include <functional>
#include <map>
#include <string>
#include <vector>

enum mock_kind { MK_NULL, MK_NUMBER, MK_STRING, MK_OBJECT, MK_EXTERNAL, MK_ARRAYBUFFER, MK_BUFFER,
	MK_FUNCTION, MK_BIGINT, MK_ARRAY };

struct napi_value__ {
	mock_kind kind = MK_NULL;
	double num = 0;
	std::string str;
	std::map<std::string, napi_value> props;
	void *ptr = nullptr;
	size_t len = 0;
	napi_callback cb = nullptr;            // native function (from napi_define_properties)
	std::function<void()> fn;              // "JS" function supplied by the harness
	std::vector<napi_value> elems;
	uint64_t big = 0;
};

napi_env mock_env_new();
bool mock_exception(napi_env env, std::string *code, std::string *msg);
napi_value mock_number(napi_env env, double d);
napi_value mock_object(napi_env env);
napi_value mock_buffer(napi_env env, void *p, size_t n);
napi_value mock_function(napi_env env, std::function<void()> f);
void mock_set(napi_value obj, const char *name, napi_value v);
napi_value mock_get(napi_value obj, const char *name);
napi_value mock_call(napi_env env, napi_value fn, std::vector<napi_value> args);
int mock_run_loop(napi_env env, int ms);
napi_value mock_load_addon(napi_env env);
