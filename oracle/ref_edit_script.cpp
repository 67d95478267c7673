// oracle/ref_edit_script.cpp — TEST INFRASTRUCTURE ONLY.
// Golden-vector generator for the alignment + canonicalisation step: links the reference's
// edlib and includes its edit_script.h, and runs the body of CEncoder::GetEditDist (encoder.cpp:1255-1283; the method is
// private, its ~20 lines are repeated here verbatim in meaning) on cases read from stdin.
// Case format (little endian): u32 kind (0 left flank, 1 right flank, 2 between anchors), u32 ref_len, u32 enc_len,
// ref_len+1 bytes (the part + the byte that follows it in the read), enc_len+1 bytes.  Output per case: u32 n, n bytes.
#include "utils.h"
#include "edit_script.h"
#include <cstdio>
#include <vector>

static EditDistRes get_edit_dist(read_view refPart, read_view encPart, uint32_t kind)
{
	if (refPart.empty() || encPart.empty())
		return get_edit_dist_on_seq_empty(refPart, encPart);
	EditDistRes ed;
	uint32_t max_symbols_for_flank = static_cast<uint32_t>(encPart.size() * 2);
	if (kind == 0)
	{
		uint32_t ref_offset;
		ed = find_edit_dist_with_edlib_ex_odwr_reverse(refPart, encPart, max_symbols_for_flank, ref_offset, EDLIB_MODE_SHW);
		refactor_edit_script(refPart.substr(ref_offset), encPart, ed.editScript);
		ed.editScript = std::string(ref_offset, 'D') + ed.editScript;
	}
	else if (kind == 1)
	{
		uint32_t tmp;
		ed = find_edit_dist_with_edlib_ex_odwr(refPart.substr(0, max_symbols_for_flank), encPart, tmp, EDLIB_MODE_SHW);
		refactor_edit_script(refPart, encPart, ed.editScript);
	}
	else
	{
		ed = find_edit_dist_with_edlib_ex(refPart, encPart);
		refactor_edit_script(refPart, encPart, ed.editScript);
	}
	return ed;
}

int main()
{
	uint32_t hdr[3];
	while (fread(hdr, 4, 3, stdin) == 3)
	{
		std::vector<uint8_t> ref(hdr[1] + 1), enc(hdr[2] + 1);
		if (fread(ref.data(), 1, ref.size(), stdin) != ref.size() || fread(enc.data(), 1, enc.size(), stdin) != enc.size()) return 1;
		EditDistRes ed = get_edit_dist(read_view(ref.data(), hdr[1]), read_view(enc.data(), hdr[2]), hdr[0]);
		uint32_t n = static_cast<uint32_t>(ed.editScript.size());
		fwrite(&n, 4, 1, stdout);
		fwrite(ed.editScript.data(), 1, n, stdout);
	}
	return 0;
}
