/* oracle/shim/boost/operators.hpp — empty stand-in: the reference config headers
 * include <boost/operators.hpp> but matrix_2d uses nothing from it. */
#ifndef GADJ_ORACLE_SHIM_BOOST_OPERATORS_HPP_
#define GADJ_ORACLE_SHIM_BOOST_OPERATORS_HPP_
#endif
