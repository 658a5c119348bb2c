// oracle/shim/htrie_map.h -- TEST INFRASTRUCTURE. Stand-in for Tessil hat-trie v0.6.0
// (reference Makefile:22-24; not vendored). The reference only uses it as a string-keyed
// associative container to collate mates (read_chimeric_alignments.cpp:16,677-680):
// insert(const char*, value) -> pair<iterator,bool>, erase(iterator), *iterator -> value.
// Iteration order is never observed, so an unordered_map is behaviourally equivalent.
#ifndef ARB_ORACLE_SHIM_HTRIE_MAP_H
#define ARB_ORACLE_SHIM_HTRIE_MAP_H
#include <string>
#include <unordered_map>
#include <utility>
namespace tsl {
template <class CharT, class T> class htrie_map {
	typedef std::unordered_map<std::basic_string<CharT>, T> map_t;
	map_t m;
public:
	class iterator {
		friend class htrie_map;
		typename map_t::iterator it;
	public:
		iterator() {}
		explicit iterator(typename map_t::iterator i): it(i) {}
		T& operator*() const { return it->second; }
		T* operator->() const { return &it->second; }
		bool operator==(const iterator& o) const { return it == o.it; }
		bool operator!=(const iterator& o) const { return it != o.it; }
	};
	std::pair<iterator,bool> insert(const CharT* key, const T& value) {
		std::pair<typename map_t::iterator,bool> r = m.insert(std::make_pair(std::basic_string<CharT>(key), value));
		return std::make_pair(iterator(r.first), r.second);
	}
	iterator erase(iterator pos) { return iterator(m.erase(pos.it)); }
	iterator end() { return iterator(m.end()); }
	size_t size() const { return m.size(); }
	bool empty() const { return m.empty(); }
};
}
#endif
