// oracle/minicv/boost -- TEST INFRASTRUCTURE.  Declarations only: the reference's DBoW2 headers name
// boost::serialization in member templates that the oracle build never instantiates.
#ifndef MINICV_BOOST_SERIALIZATION_HPP
#define MINICV_BOOST_SERIALIZATION_HPP
namespace boost {
namespace serialization {
class access;
template <class Base, class Derived>
Base& base_object(Derived& d) { return static_cast<Base&>(d); }
}  // namespace serialization
}  // namespace boost
#endif
