// TEST INFRASTRUCTURE ONLY (oracle build). Minimal std-backed stand-in for the
// boost/unordered_map.hpp facilities that the reference tools use
// (reference pins boost 1.60.0: conda/defuse/meta.yaml:12; boost is not
// installed in this image).  Written for this repo; not boost code.
//
// The reference sources say `using namespace std; using namespace boost;`
// (tools/Common.h:17-18), so the names are imported with using-declarations:
// an alias template would make `unordered_map` ambiguous.
#ifndef DFB_ORACLE_BOOST_UNORDERED_MAP_SHIM
#define DFB_ORACLE_BOOST_UNORDERED_MAP_SHIM

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <functional>
#include <map>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

namespace boost {

// Same mixing constant boost documents for hash_combine.
template <class T>
inline void hash_combine(std::size_t& seed, const T& v)
{
	seed ^= std::hash<T>()(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
}

}  // namespace boost

namespace std {

// tools/Common.h uses pair<int,int> as a hash key (IntegerPairMap, and the
// unordered_set<IntegerPair> in SplitAlignment.cpp:268,381).
template <class A, class B>
struct hash<std::pair<A, B> >
{
	std::size_t operator()(const std::pair<A, B>& p) const
	{
		std::size_t seed = 0;
		boost::hash_combine(seed, p.first);
		boost::hash_combine(seed, p.second);
		return seed;
	}
};

}  // namespace std

namespace boost {
using std::unordered_map;
using std::unordered_set;
}  // namespace boost

#endif
