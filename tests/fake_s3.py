"""A tiny S3-compatible endpoint for tests (path-style): PUT / GET / HEAD / DELETE object, ListObjectsV2 with paging, multipart
upload. Every request's AWS Signature V4 is re-derived from the raw request with the shared secret and refused (403) on mismatch;
`fail_next` injects transient 503s to exercise the client's retries."""
from __future__ import annotations

import datetime as dt
import hashlib
import hmac
import threading
import urllib.parse
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

ACCESS, SECRET, REGION = "test-access", "test-secret/with+chars", "eu-test-1"


def _sign(key: bytes, msg: str) -> bytes:
    return hmac.new(key, msg.encode(), hashlib.sha256).digest()


class FakeS3(ThreadingHTTPServer):
    daemon_threads = True

    def __init__(self, page_size: int = 1000, tls: tuple[str, str] | None = None) -> None:
        super().__init__(("127.0.0.1", 0), _Handler)
        self.scheme = "http"
        if tls is not None:          # (key file, certificate file): serve https
            import ssl

            ctx = ssl.SSLContext(ssl.PROTOCOL_TLS_SERVER)
            ctx.load_cert_chain(certfile=tls[1], keyfile=tls[0])
            self.socket = ctx.wrap_socket(self.socket, server_side=True)
            self.scheme = "https"
        self.objects: dict[str, bytes] = {}            # "bucket/key" -> data
        self.uploads: dict[str, dict[int, bytes]] = {}
        self.page_size = page_size
        self.fail_next = 0
        self.requests: list[tuple[str, str]] = []
        self.lock = threading.Lock()
        self.thread = threading.Thread(target=self.serve_forever, daemon=True)
        self.thread.start()

    @property
    def endpoint(self) -> str:
        return f"{self.scheme}://127.0.0.1:{self.server_address[1]}"

    def stop(self) -> None:
        self.shutdown()
        self.server_close()


class _Handler(BaseHTTPRequestHandler):
    protocol_version = "HTTP/1.1"

    def log_message(self, *a):  # noqa: D401 - quiet
        pass

    def _reply(self, code: int, body: bytes = b"", headers: dict[str, str] | None = None) -> None:
        self.send_response(code)
        for k, v in (headers or {}).items():
            self.send_header(k, v)
        self.send_header("Content-Length", str(len(body)))
        self.end_headers()
        if self.command != "HEAD":
            self.wfile.write(body)
        self.close_connection = self.command == "HEAD"      # a HEAD reply announces a body it does not send

    def _verify(self, body: bytes) -> bool:
        auth = self.headers.get("Authorization", "")
        try:
            cred = auth.split("Credential=")[1].split(",")[0].strip()
            signed = auth.split("SignedHeaders=")[1].split(",")[0].strip()
            sig = auth.split("Signature=")[1].strip()
            access, date, region, service, _ = cred.split("/")
        except (IndexError, ValueError):
            return False
        if access != ACCESS or region != REGION or service != "s3":
            return False
        if hashlib.sha256(body).hexdigest() != self.headers.get("x-amz-content-sha256"):
            return False
        u = urllib.parse.urlsplit(self.path)
        q = sorted(urllib.parse.parse_qsl(u.query, keep_blank_values=True))
        cq = "&".join(f"{urllib.parse.quote(k, safe='-_.~')}={urllib.parse.quote(v, safe='-_.~')}" for k, v in q)
        names = signed.split(";")
        ch = "".join(f"{n}:{' '.join((self.headers.get(n) or '').split())}\n" for n in names)
        canonical = "\n".join([self.command, u.path, cq, ch, signed, self.headers.get("x-amz-content-sha256")])
        amz = self.headers.get("x-amz-date", "")
        if abs((dt.datetime.now(dt.timezone.utc) - dt.datetime.strptime(amz, "%Y%m%dT%H%M%SZ").replace(tzinfo=dt.timezone.utc)).total_seconds()) > 900:
            return False
        scope = f"{date}/{region}/s3/aws4_request"
        to_sign = "\n".join(["AWS4-HMAC-SHA256", amz, scope, hashlib.sha256(canonical.encode()).hexdigest()])
        k = _sign(_sign(_sign(_sign(("AWS4" + SECRET).encode(), date), region), "s3"), "aws4_request")
        return hmac.compare_digest(hmac.new(k, to_sign.encode(), hashlib.sha256).hexdigest(), sig)

    def _handle(self) -> None:
        srv: FakeS3 = self.server  # type: ignore[assignment]
        n = int(self.headers.get("Content-Length") or 0)
        body = self.rfile.read(n) if n else b""
        with srv.lock:
            srv.requests.append((self.command, self.path))
            if srv.fail_next > 0:
                srv.fail_next -= 1
                return self._reply(503, b"<Error><Code>SlowDown</Code></Error>")
        if not self._verify(body):
            return self._reply(403, b"<Error><Code>SignatureDoesNotMatch</Code></Error>")
        u = urllib.parse.urlsplit(self.path)
        q = dict(urllib.parse.parse_qsl(u.query, keep_blank_values=True))
        parts = urllib.parse.unquote(u.path).lstrip("/").split("/", 1)
        bucket, key = parts[0], (parts[1] if len(parts) > 1 else "")
        full = f"{bucket}/{key}"
        with srv.lock:
            if self.command == "GET" and not key and q.get("list-type") == "2":
                pre = f"{bucket}/{q.get('prefix', '')}"
                keys = sorted(k for k in srv.objects if k.startswith(pre))
                start = int(q.get("continuation-token", "0") or 0)
                page = keys[start: start + srv.page_size]
                more = start + srv.page_size < len(keys)
                xml = ["<?xml version='1.0'?><ListBucketResult xmlns='http://s3.amazonaws.com/doc/2006-03-01/'>"]
                xml += [f"<Contents><Key>{k[len(bucket) + 1:]}</Key><Size>{len(srv.objects[k])}</Size></Contents>" for k in page]
                xml.append(f"<IsTruncated>{'true' if more else 'false'}</IsTruncated>")
                if more:
                    xml.append(f"<NextContinuationToken>{start + srv.page_size}</NextContinuationToken>")
                xml.append("</ListBucketResult>")
                return self._reply(200, "".join(xml).encode(), {"Content-Type": "application/xml"})
            if self.command == "POST" and "uploads" in q:
                uid = f"up{len(srv.uploads) + 1}"
                srv.uploads[uid] = {}
                return self._reply(200, f"<InitiateMultipartUploadResult><UploadId>{uid}</UploadId></InitiateMultipartUploadResult>".encode())
            if self.command == "PUT" and "uploadId" in q:
                srv.uploads[q["uploadId"]][int(q["partNumber"])] = body
                return self._reply(200, b"", {"ETag": f'"{hashlib.md5(body).hexdigest()}"'})
            if self.command == "POST" and "uploadId" in q:
                got = srv.uploads.pop(q["uploadId"])
                want = [int(x) for x in __import__("re").findall(r"<PartNumber>(\d+)</PartNumber>", body.decode())]
                if sorted(got) != want:
                    return self._reply(400, b"<Error><Code>InvalidPart</Code></Error>")
                srv.objects[full] = b"".join(got[i] for i in want)
                return self._reply(200, b"<CompleteMultipartUploadResult/>")
            if self.command == "DELETE" and "uploadId" in q:
                srv.uploads.pop(q["uploadId"], None)
                return self._reply(204)
            if self.command == "PUT":
                srv.objects[full] = body
                return self._reply(200, b"", {"ETag": f'"{hashlib.md5(body).hexdigest()}"'})
            if self.command in ("GET", "HEAD"):
                if full not in srv.objects:
                    return self._reply(404, b"<Error><Code>NoSuchKey</Code></Error>")
                data = srv.objects[full]
                rng = self.headers.get("Range")
                if rng and self.command == "GET":
                    lo, _, hi = rng.split("=", 1)[1].partition("-")
                    part = data[int(lo): int(hi) + 1]
                    return self._reply(206, part, {"Content-Range": f"bytes {lo}-{int(lo) + len(part) - 1}/{len(data)}"})
                return self._reply(200, data)
            if self.command == "DELETE":
                srv.objects.pop(full, None)
                return self._reply(204)
        self._reply(400, b"<Error><Code>BadRequest</Code></Error>")

    do_GET = do_PUT = do_POST = do_DELETE = do_HEAD = _handle
