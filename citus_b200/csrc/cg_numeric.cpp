/*
 * cg_numeric.cpp -- text of the coordinator-visible results.
 *
 * sum(int8) / sum(numeric) results are numeric in PostgreSQL; avg is the master-side
 * expression sum(sum) / sum(count) (planner/multi_logical_optimizer.c:1807-1830,
 * MasterAverageExpression :2302-2355) evaluated by [PG] numeric_div, whose result scale is
 * chosen by select_div_scale (numeric.c): at least NUMERIC_MIN_SIG_DIGITS = 16 significant
 * digits, counted in base-10000 digits, and never less than either input's display scale.
 * Pinned by the reference's expected/multi_tpch_query1.out (avg_qty, avg_price, avg_disc)
 * and expected/columnar_query.out:9-29 through tests/test_host_numeric.py.
 */
#include <string.h>

#include <string>

#include "cg_internal.h"

typedef unsigned __int128 u128;
typedef __int128 i128;

static std::string u128_to_string(u128 v)
{
	if (v == 0) return "0";
	char buf[48];
	int i = 47;
	buf[i] = 0;
	while (v) { buf[--i] = (char) ('0' + (int) (v % 10)); v /= 10; }
	return std::string(buf + i);
}

static int put(const std::string &s, char *buf, size_t buflen)
{
	if (s.size() + 1 > buflen) return cg_set_error(CG_EINVAL, "numeric text needs %zu bytes", s.size() + 1);
	memcpy(buf, s.c_str(), s.size() + 1);
	return CG_OK;
}

static std::string scaled_text(bool neg, const std::string &digits_in, int scale)
{
	std::string digits = digits_in;
	if (scale > 0)
	{
		if ((int) digits.size() <= scale) digits = std::string(scale + 1 - digits.size(), '0') + digits;
		digits.insert(digits.size() - scale, ".");
	}
	bool zero = digits.find_first_not_of("0.") == std::string::npos;
	return (neg && !zero ? "-" : "") + digits;
}

extern "C" int cg_numeric_out(int64_t hi, uint64_t lo, int32_t scale, char *buf, size_t buflen)
{
	if (scale < 0 || scale > 30) return cg_set_error(CG_EINVAL, "scale %d", scale);
	i128 v = ((i128) hi << 64) | (i128) (u128) lo;
	bool neg = v < 0;
	u128 mag = neg ? (u128) (-(v + 1)) + 1 : (u128) v;
	return put(scaled_text(neg, u128_to_string(mag), scale), buf, buflen);
}

/* weight (base-10000 exponent) and value of the first non-zero base-10000 digit of
 * mag * 10^-scale, as the loop at the top of select_div_scale finds them */
static void weight_firstdigit(u128 mag, int scale, int *weight, int *first)
{
	if (mag == 0) { *weight = 0; *first = 0; return; }
	int pad = (4 - scale % 4) % 4;
	for (int i = 0; i < pad; i++) mag *= 10;      /* callers keep mag < 10^34, no overflow */
	int frac_digits = (scale + pad) / 4;
	int ndig = 0, f = 0;
	while (mag) { f = (int) (mag % 10000); mag /= 10000; ndig++; }
	*weight = ndig - 1 - frac_digits;
	*first = f;
}

extern "C" int cg_numeric_div_out(int64_t hi, uint64_t lo, int32_t scale, int64_t count, char *buf, size_t buflen)
{
	if (count <= 0) return cg_set_error(CG_EINVAL, "division by %lld", (long long) count);
	if (scale < 0 || scale > 30) return cg_set_error(CG_EINVAL, "scale %d", scale);
	i128 v = ((i128) hi << 64) | (i128) (u128) lo;
	bool neg = v < 0;
	u128 mag = neg ? (u128) (-(v + 1)) + 1 : (u128) v;
	u128 limit = 1;
	for (int i = 0; i < 34; i++) limit *= 10;
	if (mag >= limit) return cg_set_error(CG_EUNSUPPORTED, "sum too large for the numeric formatter");

	int w1, f1, w2, f2;
	weight_firstdigit(mag, scale, &w1, &f1);
	weight_firstdigit((u128) count, 0, &w2, &f2);
	int qweight = w1 - w2;
	if (f1 <= f2) qweight--;
	int rscale = 16 - qweight * 4;
	if (rscale < scale) rscale = scale;
	if (rscale < 0) rscale = 0;
	if (rscale > 1000) rscale = 1000;

	/* long division: integer part, then rscale + 1 fractional digits, round half away from zero */
	u128 d = (u128) count;
	u128 scale_pow = 1;
	for (int i = 0; i < scale; i++) scale_pow *= 10;
	/* quotient of mag / (count * 10^scale) */
	u128 denom = d * scale_pow;                /* count < 2^63, scale_pow <= 10^30: fits 2^127 only for small scale */
	if (scale > 18) return cg_set_error(CG_EUNSUPPORTED, "scale %d too large for the numeric formatter", scale);
	u128 ip = mag / denom;
	u128 rem = mag % denom;
	std::string frac;
	frac.reserve(rscale + 1);
	for (int i = 0; i < rscale + 1; i++)
	{
		/* rem < denom < 2^63 * 10^18 < 2^123; rem * 10 < 2^127 */
		rem *= 10;
		int digit = (int) (rem / denom);
		rem %= denom;
		frac.push_back((char) ('0' + digit));
	}
	/* round at rscale using the extra digit */
	bool up = frac[rscale] >= '5';
	frac.resize(rscale);
	std::string ips = u128_to_string(ip);
	if (up)
	{
		int i = rscale - 1;
		while (i >= 0 && frac[i] == '9') { frac[i] = '0'; i--; }
		if (i >= 0) frac[i]++;
		else
		{
			int j = (int) ips.size() - 1;
			while (j >= 0 && ips[j] == '9') { ips[j] = '0'; j--; }
			if (j >= 0) ips[j]++;
			else ips.insert(ips.begin(), '1');
		}
	}
	std::string s = ips;
	if (rscale > 0) s += "." + frac;
	bool zero = s.find_first_not_of("0.") == std::string::npos;
	if (neg && !zero) s = "-" + s;
	return put(s, buf, buflen);
}
