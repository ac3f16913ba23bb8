#ifndef ORACLE_SHIM_LOG4CXX_PROP_H_
#define ORACLE_SHIM_LOG4CXX_PROP_H_
#include <string>
namespace log4cxx { struct PropertyConfigurator { static void configure(const std::string&) {} }; }
#endif
