#ifndef ORACLE_SHIM_LOG4CXX_BASIC_H_
#define ORACLE_SHIM_LOG4CXX_BASIC_H_
namespace log4cxx { struct BasicConfigurator { static void configure() {} }; }
#endif
