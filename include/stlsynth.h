/*
 * stlsynth.h — seeded synthetic KeyFramePack generator (host; test/bench
 * infrastructure, not part of the drop-in boundary).  See csrc/synth.cpp.
 */
#ifndef STLSYNTH_H_
#define STLSYNTH_H_
#include <stdint.h>
#include "stlcalib.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct stl_synth_cfg {
    int32_t n_kf;        /* keyframes to generate (this shard)                */
    int32_t kf_begin;    /* global index of the first one                     */
    int32_t n_kf_total;  /* keyframes of the whole sequence                   */
    int32_t beams;       /* 64  (128 for the GPR stress config)               */
    int32_t az_steps;    /* 1875 -> 120 000 rays                              */
    int32_t n_kp;        /* 2000 ORB keypoints per keyframe                   */
    int32_t n_covis;     /* 3 (num_best_covis)                                */
    int32_t width, height; /* 1241 x 376                                      */
    float fx, fy, cx, cy;  /* 718.856, 718.856, 607.1928, 185.2157            */
    double anchored_frac;  /* keypoints placed on a projected scan point      */
    double mappoint_frac;  /* anchored keypoints that carry a map point       */
    double match_frac;     /* anchored keypoints matched in a covisible KF    */
    double outlier_frac;   /* unanchored keypoints with a random match        */
    double scale_gt;       /* monocular scale s                               */
    double kf_spacing;     /* metres between keyframes                        */
    double max_range;      /* LiDAR range (m)                                 */
    double elev_top_deg, elev_bottom_deg;
    double x_gt[7];        /* GT extrinsic, Sim3 log [omega, upsilon, s]      */
    uint64_t seed;
} stl_synth_cfg_t;

typedef struct stl_synth stl_synth_t;

void stl_synth_default_cfg(stl_synth_cfg_t *cfg);
stl_synth_t *stl_synth_create(const stl_synth_cfg_t *cfg);
const stl_pack_t *stl_synth_pack(const stl_synth_t *h);
const double *stl_synth_x_gt(const stl_synth_t *h);
const double *stl_synth_Twl(const stl_synth_t *h); /* [n_kf][12] fp64 odometry */
void stl_synth_destroy(stl_synth_t *h);
void stl_synth_candidates(const double x_gt[7], uint64_t seed, int32_t B, double spread, double *out);

#ifdef __cplusplus
}
#endif
#endif
