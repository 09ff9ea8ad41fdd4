// oracle/minicv -- TEST INFRASTRUCTURE.  cv::FileStorage / cv::FileNode as the reference's
// 3rdparty/DBoW2/DBoW2/TemplatedVocabulary.h names them in its YAML save()/load() members.  Those members
// are virtual, so they are instantiated with the class, but the oracle build only calls
// loadFromTextFile() and transform(); every stub aborts if it is ever reached.
#ifndef MINICV_PERSISTENCE_STUB_HPP
#define MINICV_PERSISTENCE_STUB_HPP
#include <cstdlib>
#include <iostream>
#include <sstream>  // the real opencv2/core pulls these in; TemplatedVocabulary.h relies on it
#include <string>
namespace cv {
class FileNode {
 public:
  FileNode operator[](const char*) const { std::abort(); }
  FileNode operator[](const std::string&) const { std::abort(); }
  FileNode operator[](int) const { std::abort(); }
  size_t size() const { std::abort(); }
  operator int() const { std::abort(); }
  operator double() const { std::abort(); }
  operator float() const { std::abort(); }
  operator std::string() const { std::abort(); }
};
class FileStorage {
 public:
  enum { READ = 0, WRITE = 1 };
  FileStorage() {}
  FileStorage(const std::string&, int) { std::abort(); }
  FileStorage(const char*, int) { std::abort(); }
  bool isOpened() const { return false; }
  void release() {}
  FileNode operator[](const char*) const { std::abort(); }
  FileNode operator[](const std::string&) const { std::abort(); }
};
template <class T>
inline FileStorage& operator<<(FileStorage& fs, const T&) { std::abort(); return fs; }
}  // namespace cv
#endif
