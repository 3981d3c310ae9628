// boost/shared_ptr.hpp — stand-in: PCL 1.8's Ptr typedefs are boost::shared_ptr (reference include/ssc.h:32).
#pragma once
#include <memory>
namespace boost {
using std::make_shared;
using std::shared_ptr;
}  // namespace boost
