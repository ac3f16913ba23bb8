// Oracle build shim (test infrastructure only): stderr logger with the log4cxx macro surface the reference uses.
#ifndef ORACLE_SHIM_LOG4CXX_LOGGER_H_
#define ORACLE_SHIM_LOG4CXX_LOGGER_H_
#include <cstdlib>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
namespace log4cxx {
class Logger {
 public:
  explicit Logger(const std::string& n) : name(n) {}
  static std::shared_ptr<Logger> getLogger(const std::string& n) { return std::shared_ptr<Logger>(new Logger(n)); }
  static int level() {
    static int lvl = -1;
    if (lvl < 0) { const char* e = getenv("ORACLE_LOG_LEVEL"); lvl = e ? atoi(e) : 2; }
    return lvl;
  }
  std::string name;
};
typedef std::shared_ptr<Logger> LoggerPtr;
}  // namespace log4cxx
#define ORACLE_SHIM_LOG(lvl, tag, logger, msg) \
  do { if (log4cxx::Logger::level() <= (lvl)) { std::ostringstream _os; _os << msg; \
       std::cerr << tag << " " << (logger)->name << " - " << _os.str() << std::endl; } } while (0)
#define LOG4CXX_TRACE(logger, msg) do { } while (0)
#define LOG4CXX_DEBUG(logger, msg) do { } while (0)
#define LOG4CXX_INFO(logger, msg)  ORACLE_SHIM_LOG(2, "INFO", logger, msg)
#define LOG4CXX_WARN(logger, msg)  ORACLE_SHIM_LOG(3, "WARN", logger, msg)
#define LOG4CXX_ERROR(logger, msg) ORACLE_SHIM_LOG(4, "ERROR", logger, msg)
#define LOG4CXX_FATAL(logger, msg) ORACLE_SHIM_LOG(5, "FATAL", logger, msg)
#endif
