// TEST INFRASTRUCTURE ONLY (oracle build): see unordered_map.hpp in this directory.
#include "unordered_map.hpp"
