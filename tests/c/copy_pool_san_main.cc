/* Driver for the ThreadSanitizer build of the staging copy pool (tests/test_host_sanitizers.py). */
#include <cstdio>
extern "C" int fosphor_host_copy_selftest(int threads, unsigned long long bytes, int pieces, int rounds);
int main()
{
	int rc = fosphor_host_copy_selftest(4, (3u << 20) + 4097, 5, 10);
	rc |= fosphor_host_copy_selftest(2, 300000, 2, 10);
	rc |= fosphor_host_copy_selftest(8, 8u << 20, 8, 4);
	std::printf("copy pool self-test rc %d\n", rc);
	return rc;
}
