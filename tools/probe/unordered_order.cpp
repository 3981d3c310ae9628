// Prints what libstdc++'s std::unordered_map<int, int> does for a sequence of operations read from stdin, one per line:
//   "i <key>"  insert (no-op if present)     "e <key>"  erase (no-op if absent)     "c"  clear
// After every operation: "<bucket_count> : <keys in iteration order>".  The cluster bookkeeping of the SCV-OD path depends on this
// order (reference include/utility.h:180, src/ssc.cpp:1261 iterates cluster_set); tests/test_unordered_order.py checks a compact
// restatement of it (insert at the head of the key's bucket run, or at the front of the list when the bucket is empty; a rehash
// re-inserts every node in the old order) against this program - the specification a device-side decision table has to follow.
#include <cstdio>
#include <unordered_map>
int main() {
  std::unordered_map<int, int> m;
  char op;
  int key;
  while (scanf(" %c", &op) == 1) {
    if (op == 'c') {
      m.clear();
    } else {
      if (scanf("%d", &key) != 1) break;
      if (op == 'i') m.insert(std::make_pair(key, 0));
      if (op == 'e') m.erase(key);
    }
    printf("%zu :", m.bucket_count());
    for (auto& kv : m) printf(" %d", kv.first);
    printf("\n");
  }
  return 0;
}
