// ref_delta_roundtrip.cc — checker glue, TEST INFRASTRUCTURE ONLY.
// Parses a .delta with the REFERENCE's own reader (M_delta_stream,
// /root/reference/lib/profiles_lib/m_delta.cc:72-220) and writes it back with the
// reference's own writer (M_delta_stream_writer,
// /root/reference/lib/profiles_lib/m_delta_stream_writer.hh:55-82).  The reference sources
// are compiled where they lie (oracle/build_ref.sh); nothing of them is copied here.
// Output: the two header lines as parsed, then the re-encoded entries (the writer emits
// "1 2 3" for the three error counts, m_delta_stream_writer.hh:71).
#include <fstream>
#include <iostream>
#include <vector>

#include <m_delta.hh>
#include <m_delta_stream_writer.hh>

using namespace Para_mugsy;

int main(int argc, char **argv) {
  if (argc != 2) { std::cerr << "usage: ref_delta_roundtrip file.delta\n"; return 2; }
  std::ifstream in(argv[1]);
  if (!in) { std::cerr << "cannot open " << argv[1] << "\n"; return 2; }
  try {
    M_delta_stream ds(in);
    std::cout << ds.sequence_files().first << "\n" << ds.stream_type() << "\n";
    M_delta_stream_writer w(std::cout);
    long n = 0;
    while (M_option<M_delta_entry> de = ds.next()) { w.write(de.value()); ++n; }
    std::cerr << n << " entries\n";
  } catch (Delta_stream_parse_error const &) {
    std::cerr << "Delta_stream_parse_error\n";
    return 1;
  }
  return 0;
}
