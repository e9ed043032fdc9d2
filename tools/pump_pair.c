/*
 * pump_pair.c -- BASELINE.md row B0' (bench infrastructure): the restated plumbing of the
 * reference's data path for BASELINE configs[0], for boxes without Node.js.
 *
 * The reference moves the snapshot with two Node pipes and nothing else:
 *     zfsSend.stdout.pipe(socket)        lib/backupSender.js:179   (sender process)
 *     socket.pipe(zfsRecv.stdin)         lib/zfsClient.js:826      (receiver process)
 * i.e. read(pipe) -> write(TCP) on one thread, and read(TCP) -> write(pipe) on another, in
 * chunks of at most 64 KiB (Node's default pipe read size).  This program is exactly that and
 * nothing more -- no checksum, no codec -- so its GiB/s is the ceiling of the reference's
 * plumbing on this box, labelled "restated, not the reference" wherever it is reported.
 *
 *   pump_pair send <host> <port>     stdin  -> TCP      (sender side)
 *   pump_pair recv <port>            TCP    -> stdout   (receiver side, backlog 1 like zfsClient)
 */
#include <arpa/inet.h>
#include <errno.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <unistd.h>

#define CHUNK 65536

static int
pump(int in, int out)
{
	static char buf[CHUNK];
	for (;;) {
		ssize_t n = read(in, buf, sizeof (buf));
		if (n == 0) return (0);
		if (n < 0) { if (errno == EINTR) continue; perror("read"); return (1); }
		for (ssize_t o = 0; o < n;) {
			ssize_t w = write(out, buf + o, (size_t)(n - o));
			if (w < 0) { if (errno == EINTR) continue; perror("write"); return (1); }
			o += w;
		}
	}
}

int
main(int argc, char **argv)
{
	if (argc == 4 && strcmp(argv[1], "send") == 0) {
		struct sockaddr_in a;
		memset(&a, 0, sizeof (a));
		a.sin_family = AF_INET;
		a.sin_port = htons((unsigned short)atoi(argv[3]));
		inet_pton(AF_INET, argv[2], &a.sin_addr);
		int s = socket(AF_INET, SOCK_STREAM, 0);
		for (int tries = 0; connect(s, (struct sockaddr *)&a, sizeof (a)) != 0; tries++) {
			if (tries > 200) { perror("connect"); return (1); }
			usleep(10000);
		}
		int rc = pump(0, s);
		shutdown(s, SHUT_WR);
		close(s);
		return (rc);
	}
	if (argc == 3 && strcmp(argv[1], "recv") == 0) {
		struct sockaddr_in a;
		int one = 1;
		memset(&a, 0, sizeof (a));
		a.sin_family = AF_INET;
		a.sin_port = htons((unsigned short)atoi(argv[2]));
		a.sin_addr.s_addr = htonl(INADDR_LOOPBACK);
		int l = socket(AF_INET, SOCK_STREAM, 0);
		setsockopt(l, SOL_SOCKET, SO_REUSEADDR, &one, sizeof (one));
		if (bind(l, (struct sockaddr *)&a, sizeof (a)) != 0 || listen(l, 1) != 0) { perror("listen"); return (1); }
		int c = accept(l, NULL, NULL);
		if (c < 0) { perror("accept"); return (1); }
		int rc = pump(c, 1);
		close(c); close(l);
		return (rc);
	}
	fprintf(stderr, "usage: pump_pair send <host> <port> | pump_pair recv <port>\n");
	return (2);
}
