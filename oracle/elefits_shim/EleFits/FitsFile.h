// TEST INFRASTRUCTURE (oracle/): EleFits stand-in, see EleFits/SifFile.h.
#pragma once
#include "EleFits/SifFile.h"
