// boost shim — TEST INFRASTRUCTURE: boost::shared_ptr / make_shared / static_pointer_cast / function mapped onto the standard
// library so that the reference's vendored PCL headers compile unmodified without Boost (oracle/_ref build only).
#ifndef OPE_REFSTUB_BOOST_SHIM_
#define OPE_REFSTUB_BOOST_SHIM_
#include <functional>
#include <memory>
namespace boost {
using std::shared_ptr;
using std::make_shared;
using std::static_pointer_cast;
using std::dynamic_pointer_cast;
template <typename Sig>
class function : public std::function<Sig> {
 public:
  function() {}
  function(long) {}            // update_visualizer_ (NULL)
  function(std::nullptr_t) {}
  template <typename F> function(F f) : std::function<Sig>(f) {}
  function& operator=(const function&) = default;
  function(const function&) = default;
  bool operator!=(long) const { return static_cast<bool>(*this); }
  bool operator==(long) const { return !static_cast<bool>(*this); }
};
}  // namespace boost
#endif
