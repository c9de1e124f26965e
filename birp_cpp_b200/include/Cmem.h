// CMEM -- the analytic magnetopause / bow-shock emissivity field (Jorgensen et al. style).
//
// Public surface = the reference's include/Cmem.h:19-40.  Init() derives the model constants
// on the host exactly as the reference does (birp_cmem_init); GetSample and Camera::Render
// evaluate the field on the GPU.
#pragma once
#include <vector>

#include "../../include/birp_b200.h"
#include "Space.h"

class CMEM : public Space {
public:
    void Init(bool hack, std::vector<int> v_passed = {}, std::vector<int> b_passed = {},
              float dipole_passed = 0.f, float p0_passed = 0.786300004f, int p1_passed = 1,
              int p2_passed = 3, int p3_passed = 4, float B_passed = 2.f, float alpha_passed = 2.5f,
              float beta_passed = -1.6f, float bs_passed = 12.6400003f, float A1_passed = 7.2e-06f,
              float A2_passed = 3.5000000000000004e-06f, float ay_bs_passed = 0.f,
              float az_bs_passed = 0.f, float density_passed = 0.f);
    void Init() override;
    float GetSample(float x, float y, float z) override;

    // ---- additions (not in the reference) -----------------------------------------------
    const birp_cmem_params& Params() const { return params; }
    // evaluate the transcendental calls in fp32 on the device instead of fp64-then-round
    void SetFastMath(bool on) { fast_math = on; }
    bool FastMath() const { return fast_math; }

private:
    birp_cmem_params params{};
    bool fast_math = false;
};
