/*
 * A foreign-language client of the C ABI, in plain C99: BASELINE config #1 (bsuite.load_from_id('deep_sea/0'),
 * reset() + 1000 step() calls) driven through include/bsuite_b200.h on the explicit host path, no Python in the
 * loop.  tests/test_c_client.py compiles it with `gcc -std=c99 -pedantic -Wall -Wextra -Werror`, feeds it the
 * action mapping and the action sequence numpy produces for the reference, and compares what it prints with the
 * known answers recorded from the unmodified reference (tests/golden/known_answers.json).
 *
 *   deep_sea_client <size> <mapping.u8> <actions.i32> <num_actions_to_take> [device]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bsuite_b200.h"

static void* read_file(const char* path, size_t bytes) {
  void* data = malloc(bytes);
  FILE* fh = fopen(path, "rb");
  if (!data || !fh || fread(data, 1, bytes, fh) != bytes) {
    fprintf(stderr, "cannot read %lu bytes from %s\n", (unsigned long)bytes, path);
    exit(2);
  }
  fclose(fh);
  return data;
}

#define CHECK(call)                                                              \
  do {                                                                           \
    int32_t status__ = (call);                                                   \
    if (status__ != BSB_OK) {                                                    \
      fprintf(stderr, "%s -> %d: %s\n", #call, (int)status__, bsb_last_error()); \
      return 1;                                                                  \
    }                                                                            \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: %s size mapping.u8 actions.i32 count [device]\n", argv[0]);
    return 2;
  }
  const int size = atoi(argv[1]);
  const long count = atol(argv[4]);
  const int device = argc > 5 ? atoi(argv[5]) : BSB_DEVICE_HOST;
  unsigned char* mapping = (unsigned char*)read_file(argv[2], (size_t)size * (size_t)size);
  int32_t* actions = (int32_t*)read_file(argv[3], (size_t)count * sizeof(int32_t));

  bsb_config config;
  memset(&config, 0, sizeof(config));
  config.family = BSB_DEEP_SEA;
  config.rng_kind = BSB_RNG_MT19937; /* numpy.random.RandomState(seed), the unpatched reference's stream */
  config.size = size;
  config.deterministic = 1;
  config.unscaled_move_cost = 0.01;
  config.reward_scale = 1.0;
  config.table = mapping;
  config.table_bytes = (int64_t)size * size;

  bsb_env* env = NULL;
  CHECK(bsb_create(&config, 1, device, /*seed=*/0, /*lane_offset=*/0, &env));
  int64_t numel = 0;
  CHECK(bsb_obs_numel(env, &numel));
  if (device != BSB_DEVICE_HOST) {
    fprintf(stderr, "this client keeps its buffers in host memory: use the host path\n");
    return 2;
  }

  float* observation = (float*)calloc((size_t)numel, sizeof(float));
  double reward = 0.0;
  float discount = 0.0f;
  int32_t step_type = -1;
  bsb_outputs out;
  memset(&out, 0, sizeof(out));
  out.observation = observation;
  out.reward_f64 = &reward;
  out.discount = &discount;
  out.step_type = &step_type;

  long num_first = 0, num_last = 0, hot_cells = 0;
  double reward_sum = 0.0;
  CHECK(bsb_reset(env, &out, NULL));
  num_first += step_type == BSB_FIRST;
  for (long t = 0; t < count; ++t) {
    CHECK(bsb_step(env, &actions[t], &out, NULL));
    num_first += step_type == BSB_FIRST;
    num_last += step_type == BSB_LAST;
    if (step_type != BSB_FIRST) reward_sum += reward;
    for (int64_t k = 0; k < numel; ++k) hot_cells += observation[k] == 1.0f;
  }

  int32_t n_info = 0;
  CHECK(bsb_info_count(env, &n_info));
  printf("abi %d\n", (int)bsb_abi_version());
  printf("num_first %ld\nnum_last %ld\nhot_cells %ld\nreward_sum %.17g\n", num_first, num_last, hot_cells, reward_sum);
  for (int32_t k = 0; k < n_info; ++k) {
    double value = 0.0;
    CHECK(bsb_read_info(env, k, &value, NULL));
    printf("info %s %.17g\n", bsb_info_name(env, k), value);
  }
  int64_t steps = 0;
  CHECK(bsb_steps_done(env, &steps));
  printf("steps_done %ld\n", (long)steps);
  /* error path: a null handle is reported, not dereferenced */
  printf("null_handle_status %d\n", (int)bsb_step(NULL, &actions[0], &out, NULL));
  CHECK(bsb_destroy(env));
  free(observation);
  free(mapping);
  free(actions);
  return 0;
}
