/*
 * node_api.h -- TEST STUB, not Node's header.
 *
 * The build image has no Node.js, so js/src/binding.cc cannot be built into an addon
 * here.  This stub declares the subset of the N-API (Node-API v6) C interface the
 * binding uses, with the signatures documented for Node >= 12, so that
 * tests/test_abi.py can at least type-check the binding against include/manatee_gpu.h
 * (`g++ -fsyntax-only`).  Nothing is linked or executed.
 */
#ifndef TESTS_STUB_NODE_API_H
#define TESTS_STUB_NODE_API_H
#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct napi_env__ *napi_env;
typedef struct napi_value__ *napi_value;
typedef struct napi_callback_info__ *napi_callback_info;

typedef enum {
	napi_ok, napi_invalid_arg, napi_object_expected, napi_string_expected, napi_name_expected,
	napi_function_expected, napi_number_expected, napi_boolean_expected, napi_array_expected,
	napi_generic_failure, napi_pending_exception, napi_cancelled
} napi_status;

typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void *finalize_data, void *finalize_hint);

typedef enum {
	napi_default = 0, napi_writable = 1 << 0, napi_enumerable = 1 << 1, napi_configurable = 1 << 2,
	napi_static = 1 << 10
} napi_property_attributes;

typedef struct {
	const char *utf8name;
	napi_value name;
	napi_callback method;
	napi_callback getter;
	napi_callback setter;
	napi_value value;
	napi_property_attributes attributes;
	void *data;
} napi_property_descriptor;

napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t *argc, napi_value *argv,
    napi_value *this_arg, void **data);
napi_status napi_throw_error(napi_env env, const char *code, const char *msg);
napi_status napi_get_value_external(napi_env env, napi_value value, void **result);
napi_status napi_has_named_property(napi_env env, napi_value object, const char *utf8name, bool *result);
napi_status napi_get_named_property(napi_env env, napi_value object, const char *utf8name, napi_value *result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char *utf8name, napi_value value);
napi_status napi_get_value_double(napi_env env, napi_value value, double *result);
napi_status napi_get_value_uint32(napi_env env, napi_value value, uint32_t *result);
napi_status napi_create_external(napi_env env, void *data, napi_finalize finalize_cb, void *finalize_hint,
    napi_value *result);
napi_status napi_create_external_arraybuffer(napi_env env, void *external_data, size_t byte_length,
    napi_finalize finalize_cb, void *finalize_hint, napi_value *result);
napi_status napi_get_buffer_info(napi_env env, napi_value value, void **data, size_t *length);
napi_status napi_get_null(napi_env env, napi_value *result);
napi_status napi_create_double(napi_env env, double value, napi_value *result);
napi_status napi_create_int32(napi_env env, int32_t value, napi_value *result);
napi_status napi_create_string_utf8(napi_env env, const char *str, size_t length, napi_value *result);
napi_status napi_create_object(napi_env env, napi_value *result);
napi_status napi_create_array_with_length(napi_env env, size_t length, napi_value *result);
napi_status napi_set_element(napi_env env, napi_value object, uint32_t index, napi_value value);
napi_status napi_create_bigint_uint64(napi_env env, uint64_t value, napi_value *result);
napi_status napi_define_properties(napi_env env, napi_value object, size_t property_count,
    const napi_property_descriptor *properties);

#define NAPI_AUTO_LENGTH SIZE_MAX

typedef struct napi_threadsafe_function__ *napi_threadsafe_function;
typedef enum { napi_tsfn_release, napi_tsfn_abort } napi_threadsafe_function_release_mode;
typedef enum { napi_tsfn_nonblocking, napi_tsfn_blocking } napi_threadsafe_function_call_mode;
typedef void (*napi_threadsafe_function_call_js)(napi_env env, napi_value js_callback, void *context,
    void *data);
napi_status napi_create_threadsafe_function(napi_env env, napi_value func, napi_value async_resource,
    napi_value async_resource_name, size_t max_queue_size, size_t initial_thread_count,
    void *thread_finalize_data, napi_finalize thread_finalize_cb, void *context,
    napi_threadsafe_function_call_js call_js_cb, napi_threadsafe_function *result);
napi_status napi_call_threadsafe_function(napi_threadsafe_function func, void *data,
    napi_threadsafe_function_call_mode is_blocking);
napi_status napi_release_threadsafe_function(napi_threadsafe_function func,
    napi_threadsafe_function_release_mode mode);
napi_status napi_unref_threadsafe_function(napi_env env, napi_threadsafe_function func);

typedef napi_value (*napi_addon_register_func)(napi_env env, napi_value exports);
#ifdef __cplusplus
#define NAPI_MODULE(modname, regfunc) \
	extern "C" napi_value napi_register_module_v1(napi_env env, napi_value exports) { return regfunc(env, exports); }
#else
#define NAPI_MODULE(modname, regfunc) \
	napi_value napi_register_module_v1(napi_env env, napi_value exports) { return regfunc(env, exports); }
#endif

#ifdef __cplusplus
}
#endif
#endif
