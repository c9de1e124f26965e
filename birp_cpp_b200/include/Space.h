// Space -- what Camera integrates through.  Same two-method interface as the reference's
// include/Space.h:18-24, so existing Space subclasses (DataCube, CMEM, user classes such as the
// reference tests' CubeFake) keep working.  Camera::Render dispatches on the dynamic type:
// DataCube and CMEM are integrated entirely on the GPU; any other subclass is sampled through
// this virtual call on the host and only accumulated on the GPU.
#pragma once

class Space {
public:
    virtual void Init() = 0;
    virtual float GetSample(float x, float y, float z) = 0;
    virtual ~Space() = default;
};
