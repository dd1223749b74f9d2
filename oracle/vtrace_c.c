/* ORACLE (test infrastructure, never shipped, never the thing measured).
 *
 * Plain-C restatement of the reference's V-trace recurrence and of the
 * IMPALA loss block, used (a) as a second, independent checker next to
 * oracle/vtrace_np.py and (b) as the "port" CPU baseline that bench.py times.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the library built from this file.
 *
 * Reference lines followed (/root/reference/torchbeast/):
 *   oracle_vtrace_scan_f32   core/vtrace.py:91-139  (from_importance_weights)
 *   oracle_impala_loss_f32   core/vtrace.py:50-88 + monobeast.py:107-125,245-277
 * Arithmetic is float32 in the reference's operation order; loss sums are
 * accumulated in double (the reference uses a float32 pairwise torch.sum).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* clip < 0 or NaN means "None" (no clipping), as in include/torchbeast_b200.h */
static float clipf(float rho, float clip) { return (clip >= 0.0f && rho > clip) ? clip : rho; }

int oracle_vtrace_scan_f32(const float *log_rhos, const float *discounts, const float *rewards,
                           const float *values, const float *bootstrap, int64_t T, int64_t B,
                           float clip_rho, float clip_pg_rho, float *vs, float *pg_adv) {
  for (int64_t b = 0; b < B; ++b) {
    float acc = 0.0f;
    float vs_next = bootstrap[b];
    float v_next = bootstrap[b];
    for (int64_t t = T - 1; t >= 0; --t) {
      int64_t i = t * B + b;
      float rho = expf(log_rhos[i]);
      float c = rho < 1.0f ? rho : 1.0f;
      float delta = clipf(rho, clip_rho) * (rewards[i] + discounts[i] * v_next - values[i]);
      acc = delta + discounts[i] * c * acc;
      float v = acc + values[i];
      pg_adv[i] = clipf(rho, clip_pg_rho) * (rewards[i] + discounts[i] * vs_next - values[i]);
      vs[i] = v;
      vs_next = v;
      v_next = values[i];
    }
  }
  return 0;
}

static float log_softmax_pick(const float *row, int64_t A, int64_t a, float *lse_out) {
  float m = row[0];
  for (int64_t j = 1; j < A; ++j) m = row[j] > m ? row[j] : m;
  float s = 0.0f;
  for (int64_t j = 0; j < A; ++j) s += expf(row[j] - m);
  float lse = m + logf(s);
  if (lse_out) *lse_out = lse;
  return row[a] - lse;
}

/* Inputs are the already shifted [T,B,...] slices. rewards are raw; clip_rewards!=0 clamps to
 * [-1,1]. done is uint8/bool. losses = {pg, baseline_cost*baseline, entropy_cost*entropy}. */
int oracle_impala_loss_f32(const float *behavior_logits, const float *target_logits,
                           const int64_t *actions, const float *rewards, const uint8_t *done,
                           const float *values, const float *bootstrap, int64_t T, int64_t B,
                           int64_t A, float discounting, float baseline_cost, float entropy_cost,
                           int clip_rewards, float clip_rho, float clip_pg_rho, float *vs,
                           float *pg_adv, double *losses, float *grad_logits, float *grad_values) {
  int64_t n = T * B;
  float *log_rhos = (float *)malloc(sizeof(float) * n);
  float *disc = (float *)malloc(sizeof(float) * n);
  float *rew = (float *)malloc(sizeof(float) * n);
  if (!log_rhos || !disc || !rew) return 1;
  for (int64_t i = 0; i < n; ++i) {
    float tlp = log_softmax_pick(target_logits + i * A, A, actions[i], 0);
    float blp = log_softmax_pick(behavior_logits + i * A, A, actions[i], 0);
    log_rhos[i] = tlp - blp;
    disc[i] = done[i] ? 0.0f : discounting;
    float r = rewards[i];
    if (clip_rewards) r = r < -1.0f ? -1.0f : (r > 1.0f ? 1.0f : r);
    rew[i] = r;
  }
  oracle_vtrace_scan_f32(log_rhos, disc, rew, values, bootstrap, T, B, clip_rho, clip_pg_rho, vs, pg_adv);
  double pg = 0.0, bl = 0.0, en = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    const float *row = target_logits + i * A;
    float lse;
    float lp_a = log_softmax_pick(row, A, actions[i], &lse);
    pg += (double)(-lp_a * pg_adv[i]);
    float d = vs[i] - values[i];
    bl += 0.5 * (double)(d * d);
    float ent_row = 0.0f;
    for (int64_t j = 0; j < A; ++j) {
      float lp = row[j] - lse;
      ent_row += expf(lp) * lp;
    }
    en += (double)ent_row;
    if (grad_logits) {
      for (int64_t j = 0; j < A; ++j) {
        float lp = row[j] - lse, p = expf(lp);
        grad_logits[i * A + j] =
            pg_adv[i] * (p - (j == actions[i] ? 1.0f : 0.0f)) + entropy_cost * p * (lp - ent_row);
      }
    }
    if (grad_values) grad_values[i] = -baseline_cost * d;
  }
  losses[0] = pg;
  losses[1] = (double)baseline_cost * bl;
  losses[2] = (double)entropy_cost * en;
  free(log_rhos); free(disc); free(rew);
  return 0;
}
