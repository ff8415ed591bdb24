// Stand-in for the Boost text archive header: see ../serialization/serialization.hpp.
#pragma once
#include <boost/serialization/serialization.hpp>
