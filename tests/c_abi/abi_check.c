/* A plain C11 caller of include/ptts.h, linked against libptts_cuda.so: the boundary a Rust / C host binds
 * (INTEGRATION.md).  Checks that the header is valid C, that the structs have the layout the ctypes binding assumes,
 * and -- on a box without a GPU -- that every compute entry point refuses with PTTS_ERR_CUDA instead of falling back.
 * Exit code 0 = all good; prints one line per check. */
#include <stdio.h>
#include <string.h>

#include "ptts.h"

#define CHECK(cond)                                                 \
  do {                                                              \
    if (!(cond)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); return 1; } \
  } while (0)

int main(int argc, char** argv) {
  const int expect_no_gpu = argc > 1 && strcmp(argv[1], "--no-gpu") == 0;
  CHECK(ptts_abi_version() == PTTS_ABI_VERSION);
  CHECK(sizeof(ptts_engine_cfg) == 16 * sizeof(int32_t));
  CHECK(sizeof(ptts_tensor_desc) == 8 + 4 + 4 + 32 + 8);
  CHECK(sizeof(ptts_stream_params) == 32);
  CHECK(sizeof(ptts_segment) == 56);
  CHECK(PTTS_STEP_PCM_I16 == 4 && PTTS_SEG_TEXT == 0 && PTTS_SEG_PAUSE == 1);
  CHECK(PTTS_STEP_PCM == 1 && PTTS_STEP_AHEAD == 2 && PTTS_FRAME_OVERRUN == 2);
  printf("ok header: abi %d, cfg %zu B, tensor desc %zu B, stream params %zu B\n", ptts_abi_version(), sizeof(ptts_engine_cfg),
         sizeof(ptts_tensor_desc), sizeof(ptts_stream_params));

  /* null arguments are PTTS_ERR_INVALID everywhere, with a message */
  ptts_engine* eng = NULL;
  CHECK(ptts_engine_create(NULL, NULL, 0, &eng) == PTTS_ERR_INVALID);
  CHECK(strlen(ptts_last_error()) > 0);
  CHECK(ptts_step(NULL, NULL, 0, NULL, NULL, NULL, NULL) == PTTS_ERR_INVALID);
  CHECK(ptts_voice_from_pcm(NULL, NULL, 0, NULL) == PTTS_ERR_INVALID);
  CHECK(ptts_step_begin(NULL, NULL, 0, PTTS_STEP_PCM) == PTTS_ERR_INVALID);
  CHECK(ptts_voice_load(NULL, NULL, NULL) == PTTS_ERR_INVALID);
  CHECK(ptts_sched_create(NULL, NULL, 0, NULL) == PTTS_ERR_INVALID);
  CHECK(ptts_config_check("/nonexistent/model.yaml") == PTTS_ERR_INVALID);
  printf("ok null arguments are refused: %s\n", ptts_last_error());

  if (expect_no_gpu) {
    /* no CUDA device: creating an engine must fail loudly, there is no CPU path behind this ABI */
    ptts_engine_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.max_slots = 1; cfg.max_batch = 1; cfg.kv_capacity = 64;
    float one = 1.0f;
    ptts_tensor_desc t = {"flow_lm.bos_emb", PTTS_F32, 1, {1, 0, 0, 0}, &one};
    int32_t st = ptts_engine_create(&cfg, &t, 1, &eng);
    CHECK(st == PTTS_ERR_CUDA);
    CHECK(eng == NULL);
    printf("ok no device -> PTTS_ERR_CUDA: %s\n", ptts_last_error());
  }
  return 0;
}
