/* TEST INFRASTRUCTURE -- CPU oracle for the BiddingSimulation.step hot path.
 *
 * This is a plain-C restatement of the reference algorithm (Mikata-Project/adcraft,
 * adcraft/gymnasium_kw_env.py:160-269 -> adcraft/bidding_simulation.py:44-234 ->
 * adcraft/synthetic_kw_classes.py / synthetic_kw_helpers.py / src/lib.rs).  It is the
 * CHECKER for the CUDA path in adcraft_b200/csrc and the CPU baseline timed by
 * bench.py; it is never linked, imported or called by the product package.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py runs the unmodified
 * reference Python (from /root/reference, when present) on the same tapes, and
 * tests/golden/ holds fixtures generated that way (tests/golden/make_golden.py), plus
 * the reference notebooks' printed vectors.
 */
#ifndef ADCRAFT_ORACLE_H
#define ADCRAFT_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_SUBSTEPS 24 /* bidding_simulation.py:213 */

/* ORC_IMPLICIT_MULTI: the default ImplicitKeyword (synthetic_kw_classes.py:649-688): m ~ Binomial(
 * max_bidders, participation) bidders per lane, signed un-rounded Laplace(p1, p2) bids, cleared by
 * nth_price_auction(n=2, num_winners=1) with its zero padding for m < 3 (helpers:116-180). */
enum { ORC_IMPLICIT = 0, ORC_EXPLICIT = 1, ORC_IMPLICIT_MULTI = 2 };

/* Keyword parameters, SoA over K keywords of ONE env (gymnasium_kw_utils.py:20-28):
 * ((vol_mean, vol_std), loc|intercept, scale|slope, bctr, sctr, mean_rev, std_rev). */
typedef struct {
    int32_t kind; /* ORC_IMPLICIT / ORC_EXPLICIT */
    int32_t K;
    const double *vol_mean, *vol_std;
    const double *p1; /* implicit: Laplace loc      | explicit: impression_bid_intercept */
    const double *p2; /* implicit: Laplace scale    | explicit: impression_slope         */
    const double *ctr, *cvr;
    const double *rev_mean, *rev_std;
    const double *max_bidders, *participation; /* ORC_IMPLICIT_MULTI only (classes:659-663) */
    double impression_thresh; /* explicit only (0.05, gymnasium_kw_utils.py:81) */
} orc_keywords;

/* Replay tape of ONE env step, consumption order.  All offsets are
 * per keyword: stream[off[k] .. off[k+1]).  Streams may be longer than consumed. */
typedef struct {
    const int32_t *volume;                           /* [K] */
    const int64_t *comp_off;  const int32_t *comp_cents; /* implicit: per auction     */
    const int64_t *click_off; const double  *u_click;    /* per click slot            */
    const int64_t *conv_off;  const double  *u_conv;     /* per accepted click        */
    const int64_t *rev_off;   const int32_t *rev_cents;  /* per conversion            */
    const int32_t *impr;                             /* explicit: [K*24] binomial I; multi: [K*24] bidders m */
    const int64_t *cost_off;  const double  *cost;       /* explicit: per impression  */
    const double *comp_f64;   /* multi (shares comp_off): per auction the highest of the m bids (0 when m = 0) */
} orc_tape;

/* Optional recorder: the oracle appends what it consumed (same layout as orc_tape,
 * caller supplies capacity-sized buffers; *_n are per-keyword running counts). */
typedef struct {
    int32_t *volume;                              /* [K] */
    int64_t cap_per_kw;                           /* capacity of each per-kw stream   */
    int32_t *comp_cents; double *u_click; double *u_conv; int32_t *rev_cents;
    int32_t *impr; double *cost;                  /* impr: [K*24]                      */
    double *comp_f64;                             /* multi: per auction highest bid    */
    int32_t *n_comp, *n_click, *n_conv, *n_rev, *n_cost; /* [K] counts written         */
} orc_record;

/* Result of one env step. */
typedef struct {
    int32_t *impressions, *clicks, *conversions;  /* [K] */
    double *cost, *revenue, *profit;              /* [K] sequential f64 sums           */
    int64_t *cost_cents, *revenue_cents;          /* [K] exact integer cents (implicit)*/
    int32_t *lane_I, *lane_B, *lane_S;            /* optional [24*K] per-lane (t*K+k)  */
    double reward;
    double remaining_budget;
    int32_t lanes_run;
} orc_result;

/* ---- Philox4x32-10 and samplers (spec: DESIGN.md "Tape function") ---- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
float orc_neglog_u31(uint32_t w31);
float orc_znorm(uint32_t w);
double orc_exp(double x);
int32_t orc_laplace_cents(uint32_t w0, float loc, float scale);
int32_t orc_revenue_cents(uint32_t w, float mean, float std);
int64_t orc_volume(uint32_t w, double mean, double std);
double orc_threshold_sigmoid(double bid, double thresh, double intercept, double slope);
uint32_t orc_prob_threshold(double p);
uint64_t orc_conv_threshold(uint32_t t1, double cvr);
double orc_conv_uniform(uint32_t cc, uint32_t t1);
double orc_explicit_cost(uint32_t w3, double bid);
double orc_sum_array(const double *x, int64_t n); /* ndarray::sum order, src/lib.rs:107-111 */

/* ---- one env step ---- */
/* bids_cents: canonical integer cents (gymnasium_kw_env.py:215). budget: already round(.,2). */
/* budget_alias: 1 reproduces the reference's behaviour for ndarray budgets (the lane's
 * `budget -= cost` mutates the campaign's remaining_budget in place, bsim:102 + :225, so every
 * accepted click is charged twice); 0 is the scalar-budget behaviour (notebooks). */
int orc_step_replay(const orc_keywords *kw, const int32_t *bid_cents, double budget,
                    int budget_alias, const orc_tape *tape, orc_result *out);
int orc_step_philox(const orc_keywords *kw, const int32_t *bid_cents, double budget,
                    int budget_alias, uint64_t seed, uint32_t env_id, uint32_t step,
                    uint32_t agent, orc_result *out, orc_record *rec);

/* One bidder of a shared-auction world: draws keyed by world_id, clearing price
 * = max(sampled competitor, floor_cents[k]) where floor_cents[k] is the highest rival bid, i.e.
 * nth_price_auction(n=2, num_winners=1) on [rival bids..., competitor]
 * (synthetic_kw_helpers.py:116-180). */
int orc_step_philox_shared(const orc_keywords *kw, const int32_t *bid_cents, const int32_t *floor_cents,
                           double budget, int budget_alias, uint64_t seed, uint32_t world_id, uint32_t step,
                           orc_result *out);

/* Drift (gymnasium_kw_env.py:114-158).  coeff = [3][K] (vol, ctr, cvr); only the first
 * num_updates keywords are considered (zip truncation), masked ones updated. */
void orc_drift_apply(int32_t K, const uint8_t *mask, int32_t num_updates, const double *coeff,
                     const double *init_std, double *vol_mean, double *ctr, double *cvr);
void orc_drift_philox(int32_t K, uint64_t seed, uint32_t env_id, uint32_t step,
                      const double mag[3], double *coeff /*[3][K]*/);

/* ---- batched free-running driver (CPU baseline): E envs, shared or per-env params ---- */
typedef struct {
    int32_t kind, E, K;
    int64_t param_env_stride;     /* 0: keyword set shared by all envs, K: per-env sets */
    double *vol_mean, *vol_std, *p1, *p2, *ctr, *cvr, *rev_mean, *rev_std;
    double *max_bidders, *participation; /* ORC_IMPLICIT_MULTI, else NULL */
    double impression_thresh;
    const uint8_t *drift_mask;    /* [K] or NULL */
    double drift_mag[3];
    double *budget;               /* [E] */
    double *cum_profit;           /* [E] */
    int32_t *day;                 /* [E] */
    int32_t max_days; double loss_threshold;
    uint64_t seed; uint32_t env_base; uint32_t step;
    int32_t budget_alias;
} orc_batch;

/* bids [E*K] f64 dollars (canonicalised inside), outputs [E*K] / [E]. Returns 0. */
int orc_batch_step(orc_batch *b, const double *bids, int32_t *impressions, int32_t *clicks,
                   int32_t *conversions, double *cost, double *revenue, double *reward,
                   uint8_t *terminated, uint8_t *truncated, int n_threads);

int32_t orc_bid_to_cents(double bid);

#ifdef __cplusplus
}
#endif
#endif
