// CMEM on top of the C ABI: constants on the host (birp_cmem_init, bit-identical to the
// reference's CMEM::Init), samples on the GPU (birp_sample_cmem / birp_render_cmem).
#include "Cmem.h"

#include <stdexcept>
#include <string>

void CMEM::Init() { this->Init(true); }

void CMEM::Init(bool, std::vector<int> v_passed, std::vector<int> b_passed, float dipole_passed,
                float p0_passed, int p1_passed, int p2_passed, int p3_passed, float B_passed,
                float alpha_passed, float beta_passed, float bs_passed, float A1_passed, float A2_passed,
                float ay_bs_passed, float az_bs_passed, float density_passed) {
    // the reference indexes v_passed[0..2] without checking; a short vector is a caller error
    if ((!v_passed.empty() && v_passed.size() < 3) || (!b_passed.empty() && b_passed.size() < 3))
        throw std::invalid_argument("CMEM::Init: v and b need three components");
    int rc = birp_cmem_init(!v_passed.empty(), v_passed.data(), !b_passed.empty(), b_passed.data(),
                            dipole_passed, p0_passed, p1_passed, p2_passed, p3_passed, B_passed, alpha_passed,
                            beta_passed, bs_passed, A1_passed, A2_passed, ay_bs_passed, az_bs_passed,
                            density_passed, &params);
    if (rc != BIRP_OK) throw std::runtime_error(std::string("CMEM::Init: ") + birp_last_error());
}

float CMEM::GetSample(float x, float y, float z) {
    const float xyz[3] = {x, y, z};
    float out = 0.f;
    int rc = birp_sample_cmem(&params, xyz, 1, fast_math ? BIRP_CMEM_FAST : 0, &out);
    if (rc != BIRP_OK) throw std::runtime_error(std::string("CMEM::GetSample: ") + birp_last_error());
    return out;
}
