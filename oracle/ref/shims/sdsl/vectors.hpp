#pragma once
#include <sdsl/int_vector.hpp>
