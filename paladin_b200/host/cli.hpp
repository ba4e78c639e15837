// cli.hpp -- "--key=value" command line handling with the reference's semantics
// (reference src/command-line-parser.hpp:39-123): the first two characters of every argument are dropped,
// the remainder is split at '=', a key without value maps to "n/a", the first occurrence of a key wins and
// unknown keys are only warned about by the caller.
#ifndef PALADIN_GPU_HOST_CLI_HPP_
#define PALADIN_GPU_HOST_CLI_HPP_

#include <algorithm>
#include <string>
#include <vector>

namespace paladin {

// tokens between delimiters, empty tokens dropped (so repeated delimiters collapse)
inline std::vector<std::string> split_string( const std::string& text, char delimiter = ' ' ) {
  std::vector<std::string> out;
  std::size_t from = 0;
  while ( from <= text.size() ) {
    std::size_t to = text.find( delimiter, from );
    if ( to == std::string::npos ) to = text.size();
    if ( to > from ) out.emplace_back( text, from, to - from );
    from = to + 1;
  }
  return out;
}

class CommandLineParser {
  std::vector<std::pair<std::string, std::string>> options_;

 public:
  CommandLineParser( int argc, char* argv[] ) {
    for ( int a = 1; a < argc; ++a ) {
      std::string arg( argv[a] );
      arg.erase( 0, std::min<std::size_t>( 2, arg.size() ) );
      const std::vector<std::string> kv = split_string( arg, '=' );
      if ( kv.empty() ) {
        options_.emplace_back( std::string(), "n/a" );
        continue;
      }
      options_.emplace_back( kv[0], kv.size() > 1 ? kv[1] : std::string( "n/a" ) );
    }
  }

  bool checkExists( const std::string& key ) const {
    for ( const auto& o : options_ )
      if ( o.first == key ) return true;
    return false;
  }

  std::string getValue( const std::string& key, const std::string& fallback ) const {
    for ( const auto& o : options_ )
      if ( o.first == key ) return o.second;
    return fallback;
  }

  std::vector<std::string> get_keys() const {
    std::vector<std::string> keys;
    for ( const auto& o : options_ ) keys.push_back( o.first );
    return keys;
  }
};

}  // namespace paladin

#endif
