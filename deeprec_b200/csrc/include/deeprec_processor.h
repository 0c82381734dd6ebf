/* deeprec_b200 serving C ABI (the reference: serving/processor/serving/processor.h -- initialize / process / batch_process /
 * get_serving_model_info; return code 200 = OK, 500 = error).
 *
 * Two runtimes export it:
 *   libdeeprec_cuda.so   GPU runtime            initialize, process, batch_process, get_serving_model_info, dr_serving_release, dr_serving_free
 *   libdeeprec_host.so   CPU runtime (no GPU)   the same functions with a dr_cpu_ prefix (both libraries may live in one process)
 *
 * model_config is the JSON ModelConfig: session_num, select_session_policy ("RR" | "MOD"), max_batch, checkpoint_dir,
 * model_update_interval_ms, warmup_file_name, timeline_{start_step,interval_step,trace_count,path}; GPU: gpu_id, mlp_dtype ("bf16" | "fp8"),
 * delta_extra_rows; CPU: intra_op_parallelism_threads, feature_store_type ("local" | "redis"), redis_url, redis_password, redis_db_idx,
 * redis_prefix, redis_timeout_ms.
 *
 * A request is either the compact format (struct dr_wire_request below, then dense[batch][num_dense] float32, then ids[num_sparse][batch]
 * int64) or a protobuf tensorflow.eas.PredictRequest (inputs "dense" + "ids", or one input per feature I1.. / C1..).  The reply uses the
 * same encoding as the request.  Output buffers are malloc'ed by the runtime: release them with dr_serving_free / dr_cpu_serving_free.
 */
#ifndef DEEPREC_PROCESSOR_H_
#define DEEPREC_PROCESSOR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#pragma pack(push, 1)
struct dr_wire_request { uint32_t magic /* 0x51525244 "DRRQ" */, version /* 1 */, batch, num_dense, num_sparse, reserved; };
struct dr_wire_response { uint32_t magic /* 0x53525244 "DRRS" */, batch, status, reserved; int64_t model_version; /* then float32 probabilities[batch] */ };
#pragma pack(pop)

/* ---- GPU runtime (libdeeprec_cuda.so) ---- */
void* initialize(const char* model_entry, const char* model_config, int* state);
int process(void* model_buf, const void* input_data, int input_size, void** output_data, int* output_size);
/* input_size[0] = number of requests n, input_size[1..n] = their sizes */
int batch_process(void* model_buf, const void* input_data[], int* input_size, void* output_data[], int* output_size);
int get_serving_model_info(void* model_buf, void** output_data, int* output_size);   /* JSON */
void dr_serving_release(void* model_buf);
void dr_serving_free(void* p);

/* ---- CPU runtime (libdeeprec_host.so) ---- */
void* dr_cpu_initialize(const char* model_entry, const char* model_config, int* state);
int dr_cpu_process(void* model_buf, const void* input_data, int input_size, void** output_data, int* output_size);
int dr_cpu_batch_process(void* model_buf, const void* input_data[], int* input_size, void* output_data[], int* output_size);
int dr_cpu_get_serving_model_info(void* model_buf, void** output_data, int* output_size);
void dr_cpu_serving_release(void* model_buf);
void dr_cpu_serving_free(void* p);

/* ---- protobuf helpers for C clients (libdeeprec_host.so); buffers returned through `out` are released with dr_pb_free ---- */
/* per_feature = 0: inputs {"dense": float[B, nd], "ids": int64[ns, B]};  1: I1..I<nd> and C1..C<ns> */
int dr_pb_encode_request(const float* dense, const int64_t* ids, int64_t B, int nd, int ns, int per_feature, const char* signature,
                         const char* output_filter, void** out, int64_t* out_n);
/* returns the number of probabilities (copied up to cap), or -1 */
int64_t dr_pb_decode_response(const void* pb, int64_t n, float* probs, int64_t cap, int64_t* model_version);
int64_t dr_pb_response_cols(const void* pb, int64_t n);   /* probabilities per row (multi-task models: > 1, sample-major) */
const char* dr_pb_last_error(void);
void dr_pb_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* DEEPREC_PROCESSOR_H_ */
