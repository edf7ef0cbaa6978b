// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for the two boost string
// algorithms the reference tools use: split(..., is_any_of(...)) with empty
// tokens kept (no token compression), e.g. tools/localalign.cpp:67,
// tools/AlignmentStream.cpp:58,93.  Written for this repo; not boost code.
#ifndef DFB_ORACLE_BOOST_ALGORITHM_STRING_SHIM
#define DFB_ORACLE_BOOST_ALGORITHM_STRING_SHIM

#include <string>
#include <vector>

namespace boost {

struct shim_any_of
{
	std::string chars;
	bool operator()(char c) const { return chars.find(c) != std::string::npos; }
};

inline shim_any_of is_any_of(const std::string& chars)
{
	shim_any_of p;
	p.chars = chars;
	return p;
}

template <class Pred>
inline std::vector<std::string>& split(std::vector<std::string>& out, const std::string& input, Pred pred)
{
	out.clear();
	std::string token;
	for (std::string::size_type k = 0; k < input.size(); k++)
	{
		if (pred(input[k]))
		{
			out.push_back(token);
			token.clear();
		}
		else
		{
			token.push_back(input[k]);
		}
	}
	out.push_back(token);
	return out;
}

}  // namespace boost

#endif
