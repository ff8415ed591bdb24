// Minimal stand-in for <boost/serialization/serialization.hpp>.
// Boost is not installed in this image.  The reference's hot-path classes only
// name boost::serialization::access in a friend declaration and keep their
// serialize() members as never-instantiated templates (Coord.h:71-79,
// Node.h:53-65), so a forward declaration is all that is needed to compile the
// reference's arithmetic files unmodified.  The two std headers arrive
// transitively with real Boost and the reference relies on that.
#pragma once
#include <cstdint>
#include <limits>
namespace boost { namespace serialization { class access; } }
