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
  MPPIB_DYN_QUADROTOR = 4,         /* dynamics/quadrotor/quadrotor_dynamics.cuh          S13 C4 O13 */
  MPPIB_DYN_COUNT
};

enum mppib_cost_id
{
  MPPIB_COST_CARTPOLE_QUADRATIC = 0, /* cost_functions/cartpole/cartpole_quadratic_cost.cuh */
  MPPIB_COST_DI_CIRCLE = 1,          /* cost_functions/double_integrator/double_integrator_circle_cost.cuh */
  MPPIB_COST_AR_STANDARD = 2,        /* cost_functions/autorally/ar_standard_cost.cuh */
  MPPIB_COST_RACER_QUADRATIC = 3,    /* ours: quadratic on speed / yaw, documented in DESIGN.md */
  MPPIB_COST_QUADROTOR_QUADRATIC = 4, /* cost_functions/quadrotor/quadrotor_quadratic_cost.cuh */
  MPPIB_COST_COUNT
};

enum mppib_sampler_id
{
  MPPIB_SAMPLER_GAUSSIAN = 0,     /* sampling_distributions/gaussian/gaussian.cuh */
  MPPIB_SAMPLER_COLORED_NOISE = 1, /* sampling_distributions/colored_noise/colored_noise.cuh */
  MPPIB_SAMPLER_NLN = 2            /* sampling_distributions/nln/nln.cuh: normal x log-normal noise, GaussianParams */
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

/* RacerDubinsElevationLSTMSteering (dynamics/racer_dubins/racer_dubins_elevation_lstm_steering.cuh): the parametric
 * fields of RacerDubinsParams (racer_dubins.cuh:78-104) + RacerDubinsElevationParams (racer_dubins_elevation.cuh:47-59).
 * The LSTM architecture is a constructor argument in the reference (lstm_steering.cu:11-22) and therefore travels in
 * mppib_desc.model_dims = { hidden_dim H, head hidden width L1 } (LSTM input dim is 4, head layers {H+4, L1, 1});
 * the weights travel as MPPIB_BLOB_LSTM_WEIGHTS in the reference's packed layouts:
 *   LSTM  W_im W_fm W_om W_cm (H x H row-major each) | W_ii W_fi W_oi W_ci (H x 4) | b_i b_f b_o b_c (H) |
 *         initial_hidden (H) | initial_cell (H)                               (utils/nn_helpers/lstm_helper.cu:72-88)
 *   head  per layer W (row-major out x in) then b                              (utils/nn_helpers/fnn_helper.cu:176-183)
 * Elevation map: not built yet (flat terrain == TwoDTextureHelper::checkTextureUse(0) false, racer_dubins.cu:427-432). */
#define MPPIB_RACER_LSTM_INPUT_DIM 4
#define MPPIB_RACER_LSTM_NUM_PARAMS(H, L1) \
  (4 * (H) * (H) + 4 * (H) * 4 + 6 * (H) + ((H) + 4) * (L1) + (L1) + (L1) + 1)
typedef struct mppib_racer_lstm_dyn_params
{
  mppib_control_limits lim;
  float c_t[3];                    /* 1.3, 2.6, 3.9 */
  float c_b[3];                    /* 2.5, 3.5, 4.5 */
  float c_v[3];                    /* 3.7, 4.7, 5.7 */
  float c_0;                       /* 4.9 */
  float steering_constant;         /* 0.6 */
  float steer_command_angle_scale; /* 5 */
  float steer_angle_scale;         /* -9.1 */
  float max_steer_angle;           /* 0.5 */
  float max_steer_rate;            /* 5 */
  float steer_accel_constant;      /* 12.1 */
  float steer_accel_drag_constant; /* 1.0 */
  float brake_delay_constant;      /* 6.6 */
  float brake_delay_constant_neg;  /* 8.2 */
  float max_brake_rate_neg;        /* 0.9 */
  float max_brake_rate_pos;        /* 0.33 */
  float wheel_base;                /* 0.3 */
  float low_min_throttle;          /* 0.13 */
  float gravity;                   /* -9.81 */
  int gear_sign;                   /* 1 */
  float clamp_ax;                  /* 5.5 */
  float K_x, K_y, K_yaw, K_vel_x;  /* 1 */
  float Q_x_acc;                   /* 1 */
  float Q_x_v[3];                  /* 41.74219, -0.8187027, -2.2131343 */
  float Q_y_f;                     /* 0.1 */
  float Q_omega_v;                 /* 0.001 */
  float Q_omega_steering;          /* 0 */
} mppib_racer_lstm_dyn_params;

/* Elevation map of the RACER models (utils/texture_helpers/texture_helper.cuh:11-56 TextureParams + two_d_texture_helper.cu):
 * MPPIB_BLOB_ELEVATION_MAP = this header followed by width * height floats, row-major (value at row i, column j =
 * data[i * width + j]: TwoDTextureHelper's cpu_values_ layout). Clamp addressing, bilinear filtering, normalised
 * coordinates (the TextureParams defaults). `use` is enableTexture / checkTextureUse (texture_helper.cu): with use == 0 the
 * model settles on flat ground (racer_dubins.cu:427-432). */
typedef struct mppib_elevation_map_header
{
  int width, height;   /* cudaExtent: width = x cells, height = y cells (both >= 2) */
  float origin[3];     /* TextureParams::origin */
  float rotations[9];  /* TextureParams::rotations[3], row-major: map = R (world - origin) */
  float resolution[3]; /* metres per cell */
  int use;
} mppib_elevation_map_header;

/* QuadrotorDynamics (dynamics/quadrotor/quadrotor_dynamics.cuh:9-63). State POS(3) VEL(3) QUAT_W..Z(4) ANG_VEL(3);
 * controls ANG_RATE_X/Y/Z, THRUST. The default constructor's thrust range [0, 36] and zero_control[3] = GRAVITY
 * (quadrotor_dynamics.cu:11-19) are set by the host-side mirror, they are not part of the params struct. */
#define MPPIB_GRAVITY 9.81f /* utils/math_utils.h:45 */
typedef struct mppib_quadrotor_dyn_params
{
  mppib_control_limits lim;
  float tau_roll;  /* 0.25 */
  float tau_pitch; /* 0.25 */
  float tau_yaw;   /* 0.25 */
  float mass;      /* 1 kg */
} mppib_quadrotor_dyn_params;

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

/* Quadratic tracking cost on the RACER output vector (ours: the RACER cost classes are not in the reference tree,
 * SURVEY §8d C5). cost = speed_coeff (y[VEL_B_X] - desired_speed)^2 + yaw_coeff angdist(y[YAW], desired_yaw)^2
 *                      + lateral_coeff (y[POS_I_Y] - desired_y)^2 + steer_coeff y[STEER_ANGLE]^2, times discount^t;
 * terminal cost 0. */
typedef struct mppib_racer_quadratic_cost_params
{
  float control_cost_coeff[MPPIB_MAX_CONTROL_DIM];
  float discount;      /* 1.0 */
  float desired_speed; /* 5.0 */
  float speed_coeff;   /* 4.0 */
  float desired_yaw;   /* 0.0 */
  float yaw_coeff;     /* 20.0 */
  float desired_y;     /* 0.0 */
  float lateral_coeff; /* 2.0 */
  float steer_coeff;   /* 1.0 */
} mppib_racer_quadratic_cost_params;

/* QuadrotorQuadraticCost (cost_functions/quadrotor/quadrotor_quadratic_cost.cuh:9-66) */
typedef struct mppib_quadrotor_cost_params
{
  float control_cost_coeff[MPPIB_MAX_CONTROL_DIM]; /* 2, 2, 2, 2; unused on device (cost.cuh:205-208) */
  float discount;                                  /* CostParams default 1.0 (unused by this cost) */
  float s_goal[13];                                /* x(3) v(3) q(4: 1,0,0,0) w(3) */
  float x_coeff;                                   /* 1 */
  float v_coeff;                                   /* 1 */
  int use_euler;                                   /* true */
  float q_coeff;                                   /* 1 */
  float roll_coeff;                                /* 1 */
  float pitch_coeff;                               /* 1 */
  float yaw_coeff;                                 /* 1 */
  float w_coeff;                                   /* 1 */
  float terminal_cost_coeff;                       /* 0 */
} mppib_quadrotor_cost_params;
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
