// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for boost::lexical_cast as the
// reference tools call it (e.g. tools/SplitAlignment.cpp:259,282; tools/Parsers.cpp:242-251).
// Written for this repo; not boost code.
#ifndef DFB_ORACLE_BOOST_LEXICAL_CAST_SHIM
#define DFB_ORACLE_BOOST_LEXICAL_CAST_SHIM

#include <sstream>
#include <string>
#include <typeinfo>

namespace boost {

class bad_lexical_cast : public std::bad_cast
{
public:
	virtual const char* what() const throw() { return "bad lexical cast"; }
};

// The reference names this exact nested type in its catch clauses
// (tools/Parsers.cpp:76,161,256; tools/ExonRegions.cpp:61).
namespace exception_detail {
template <class T> struct error_info_injector : public T {};
template <class T> struct clone_impl : public T {};
}  // namespace exception_detail

namespace shim_detail {
typedef exception_detail::clone_impl<exception_detail::error_info_injector<bad_lexical_cast> > thrown_type;

template <class Target>
struct caster
{
	template <class Source>
	static Target apply(const Source& src)
	{
		std::stringstream ss;
		Target result;
		// whole-token conversion: trailing characters are an error, like boost
		if (!(ss << src) || !(ss >> result) || !(ss >> std::ws).eof())
		{
			throw thrown_type();
		}
		return result;
	}
};

template <>
struct caster<std::string>
{
	template <class Source>
	static std::string apply(const Source& src)
	{
		std::stringstream ss;
		ss << src;
		return ss.str();
	}
};
}  // namespace shim_detail

template <class Target, class Source>
inline Target lexical_cast(const Source& src)
{
	return shim_detail::caster<Target>::apply(src);
}

}  // namespace boost

#endif
