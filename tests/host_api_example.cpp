// The reference's API example (src/API_example/api_example.cpp) written against include/colord_b200_api.h: prints the archive's
// info to stderr and its records to stdout.  Compiled by tests/test_host_decode.py (host only).
#include "../include/colord_b200_api.h"
#include <iostream>

int main(int argc, char** argv)
{
	if (argc < 2) { std::cerr << "Usage: " << argv[0] << " <colord-b200 archive>\n"; return 1; }
	try {
		colord_b200::DecompressionStream stream(argv[1]);
		const auto info = stream.GetInfo();
		std::cerr << "Database info:\n\n";
		info.ToOstream(std::cerr);
		while (auto x = stream.NextRecord()) {
			if (info.isFastq) std::cout << "@" << x.ReadHeader() << "\n" << x.Read() << "\n" << "+" << x.QualHeader() << "\n" << x.Qual() << "\n";
			else std::cout << ">" << x.ReadHeader() << "\n" << x.Read() << "\n";
		}
	} catch (const std::exception& e) {
		std::cerr << "Error: " << e.what() << "\n";
		return 1;
	}
	return 0;
}
