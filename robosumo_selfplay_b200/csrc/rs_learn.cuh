// rs_learn.cuh -- learner-side kernels (policy inference, V-trace, PPO update); filled in below
#pragma once
