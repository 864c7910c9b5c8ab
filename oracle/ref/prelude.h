// Force-included before every reference TU: the reference relies on transitive includes
// that newer libstdc++ no longer provides.
#pragma once
#ifdef __cplusplus
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <string>
#include <vector>
#include <map>
#include <set>
#include <unordered_map>
#include <random>
#include <mutex>
#include <thread>
#include <condition_variable>
#include <iostream>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <algorithm>
#include <chrono>
#include <ctime>
#endif
