// ORACLE — TEST INFRASTRUCTURE ONLY.  NOT Sophus: see se3.hpp in this directory (SO3d lives there).
#pragma once
#include "se3.hpp"
