/*
 * mppi_b200/params.h — plain-C parameter blobs that cross the C-ABI (include/mppi_b200.h).
 *
 * The reference keeps every plugin's parameters in a POD struct that it memcpy's to the device
 * (include/mppi/utils/managed.cuh:121-131). Templates cannot cross a C ABI, so the same information
 * crosses it here as fixed-layout C structs selected by a plugin id. Each struct cites the reference
 * struct whose fields it carries. All floats are IEEE binary32, all ints 32 bit, no padding surprises
 * (only 4-byte members).
 */
#ifndef MPPI_B200_PARAMS_H_
#define MPPI_B200_PARAMS_H_

#ifdef __cplusplus
extern "C" {
#endif

#define MPPIB_MAX_CONTROL_DIM 4
#define MPPIB_MAX_DISTRIBUTIONS 2 /* sampling_distributions/gaussian/gaussian.cuh:20-25 (MAX_DISTRIBUTIONS_T = 2) */

/* ---- plugin ids -------------------------------------------------------------------------------- */
enum mppib_dynamics_id
{
  MPPIB_DYN_CARTPOLE = 0,          /* dynamics/cartpole/cartpole_dynamics.cuh            S4 C1 O4 */
  MPPIB_DYN_DOUBLE_INTEGRATOR = 1, /* dynamics/double_integrator/di_dynamics.cuh         S4 C2 O4 */
  MPPIB_DYN_AUTORALLY_NN = 2,      /* dynamics/autorally/ar_nn_model.cuh NeuralNetModel<7,2,3>  S7 C2 O8 */
  MPPIB_DYN_RACER_LSTM = 3,        /* dynamics/racer_dubins/racer_dubins_elevation_lstm_steering.cuh S19 C2 O28 */
  MPPIB_DYN_COUNT
};

enum mppib_cost_id
{
  MPPIB_COST_CARTPOLE_QUADRATIC = 0, /* cost_functions/cartpole/cartpole_quadratic_cost.cuh */
  MPPIB_COST_DI_CIRCLE = 1,          /* cost_functions/double_integrator/double_integrator_circle_cost.cuh */
  MPPIB_COST_AR_STANDARD = 2,        /* cost_functions/autorally/ar_standard_cost.cuh */
  MPPIB_COST_RACER_QUADRATIC = 3,    /* ours: quadratic on speed / yaw, documented in DESIGN.md */
  MPPIB_COST_COUNT
};

enum mppib_sampler_id
{
  MPPIB_SAMPLER_GAUSSIAN = 0,     /* sampling_distributions/gaussian/gaussian.cuh */
  MPPIB_SAMPLER_COLORED_NOISE = 1 /* sampling_distributions/colored_noise/colored_noise.cuh */
};

/* ---- Dynamics base: control limits (dynamics/dynamics.cuh:133,511-512) ------------------------- */
typedef struct mppib_control_limits
{
  float rng_lo[MPPIB_MAX_CONTROL_DIM];   /* control_rngs_[i].x, default -FLT_MAX (dynamics.cuh:103) */
  float rng_hi[MPPIB_MAX_CONTROL_DIM];   /* control_rngs_[i].y, default +FLT_MAX (dynamics.cuh:104) */
  float deadband[MPPIB_MAX_CONTROL_DIM]; /* control_deadband_, default 0 */
  float zero_control[MPPIB_MAX_CONTROL_DIM]; /* zero_control_, default 0 */
} mppib_control_limits;

/* ---- Dynamics parameter blobs ------------------------------------------------------------------- */
typedef struct mppib_cartpole_dyn_params /* dynamics/cartpole/cartpole_dynamics.cuh:27-37 */
{
  mppib_control_limits lim;
  float cart_mass;   /* default 1 */
  float pole_mass;   /* default 1 */
  float pole_length; /* default 1 */
  float gravity;     /* gravity_ = 9.81 (cartpole_dynamics.cuh:101) */
} mppib_cartpole_dyn_params;

typedef struct mppib_di_dyn_params /* dynamics/double_integrator/di_dynamics.cuh:9-25 */
{
  mppib_control_limits lim;
  float system_noise; /* host-side disturbance only; unused on the rollout path */
} mppib_di_dyn_params;

/* NeuralNetModel<7,2,3>: limits here; the 6-32-32-4 weights travel as a separate blob
 * (MPPIB_BLOB_NN_WEIGHTS) in the reference's packed layout: per layer W (row-major out x in) then b
 * (utils/nn_helpers/fnn_helper.cu:176-183). */
#define MPPIB_AR_NN_NUM_PARAMS 1412 /* (6+1)*32 + (32+1)*32 + (32+1)*4 */
typedef struct mppib_ar_nn_dyn_params
{
  mppib_control_limits lim;
} mppib_ar_nn_dyn_params;

/* ---- Cost parameter blobs ----------------------------------------------------------------------- */
typedef struct mppib_cartpole_cost_params /* cost_functions/cartpole/cartpole_quadratic_cost.cuh:10-23 */
{
  float control_cost_coeff[MPPIB_MAX_CONTROL_DIM]; /* CostParams<1> (cost.cuh:17-31); unused on device (cost.cuh:205-208) */
  float discount;
  float cart_position_coeff;         /* 1000 */
  float cart_velocity_coeff;         /* 100 */
  float pole_angle_coeff;            /* 2000 */
  float pole_angular_velocity_coeff; /* 100 */
  float terminal_cost_coeff;         /* 0 */
  float desired_terminal_state[4];   /* {0,0,pi,0} */
} mppib_cartpole_cost_params;

typedef struct mppib_di_circle_cost_params /* cost_functions/double_integrator/double_integrator_circle_cost.cuh:8-23 */
{
  float control_cost_coeff[MPPIB_MAX_CONTROL_DIM];
  float discount;                 /* 1.0 */
  float velocity_cost;            /* 1 */
  float crash_cost;               /* 1000 */
  float velocity_desired;         /* 2 */
  float inner_path_radius2;       /* 1.875^2 */
  float outer_path_radius2;       /* 2.125^2 */
  float angular_momentum_desired; /* 2*velocity_desired */
} mppib_di_circle_cost_params;

typedef struct mppib_ar_standard_cost_params /* cost_functions/autorally/ar_standard_cost.cuh:14-41 */
{
  float control_cost_coeff[MPPIB_MAX_CONTROL_DIM];
  float discount;           /* CostParams default 1.0 */
  float desired_speed;      /* 6.0 */
  float speed_coeff;        /* 4.25 */
  float track_coeff;        /* 200 */
  float max_slip_ang;       /* 1.25 */
  float slip_coeff;         /* 10 */
  float track_slop;         /* 0 */
  float crash_coeff;        /* 10000 */
  float boundary_threshold; /* 0.65 */
  int grid_res;             /* 10 (unused on the path) */
  float r_c1[3];            /* R matrix col 1 */
  float r_c2[3];            /* R matrix col 2 */
  float trs[3];             /* translation */
  int l1_cost;              /* ARStandardCostImpl::l1_cost_ (ar_standard_cost.cuh), default 0 */
  float front_d;            /* FRONT_D = 0.5  (ar_standard_cost.cuh) */
  float back_d;             /* BACK_D = -0.5 */
  int map_width;            /* texture width  (costmap travels as MPPIB_BLOB_COSTMAP, float4 per texel) */
  int map_height;           /* texture height */
} mppib_ar_standard_cost_params;

/* ---- Sampler parameter blob (sampling_distribution.cuh:14-29, gaussian.cuh:21-61) --------------- */
typedef struct mppib_gaussian_params
{
  float std_dev[MPPIB_MAX_CONTROL_DIM * MPPIB_MAX_DISTRIBUTIONS]; /* [d][c] with stride CONTROL_DIM of the plugin */
  float control_cost_coeff[MPPIB_MAX_CONTROL_DIM];                /* default 0 (gaussian.cuh:26) */
  float pure_noise_trajectories_percentage;                       /* 0.01 */
  float std_dev_decay;                                            /* 1.0 */
  int sum_strides;                          /* 32; kept for API parity, the B200 reduction does not use it */
  int use_same_noise_for_all_distributions; /* 1 (sampling_distribution.cuh:20) */
  /* ColoredNoise extras (colored_noise.cuh:41-60); ignored by the Gaussian sampler */
  float exponents[MPPIB_MAX_CONTROL_DIM * MPPIB_MAX_DISTRIBUTIONS]; /* default 0 == white */
  float offset_decay_rate;                                          /* 0.97 */
  float fmin;                                                       /* 0.0 */
} mppib_gaussian_params;

#ifdef __cplusplus
}
#endif
#endif /* MPPI_B200_PARAMS_H_ */
