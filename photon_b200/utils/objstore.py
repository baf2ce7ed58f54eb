"""Object stores behind the server checkpoints and the ``s3`` communication stack.

The reference reaches S3 through Composer's ``RemoteUploaderDownloader`` (boto3 underneath; ref: photon/server/s3_utils.py:215-330,
812-864, configured by ``s3_comm_config`` + the ``S3_ENDPOINT_URL`` / ``AWS_*`` environment, ref: photon/conf/base_schema.py:265-277).
boto3 is not a dependency here: :class:`S3ObjectStore` speaks the S3 REST API directly (AWS Signature Version 4 over ``http.client``,
path-style addressing so MinIO / Ceph / AWS all work): PUT / GET / HEAD / DELETE object, ListObjectsV2 with continuation, multipart
upload for large files (parts in flight in parallel, each retried on its own), parallel ranged downloads, bounded retries
with back-off on 5xx / connection errors. :class:`DirObjectStore` is the same interface over a directory (the offline default).

Keys are ``/``-separated strings relative to the store (bucket + optional prefix).
"""
from __future__ import annotations

import datetime as _dt
import hashlib
import hmac
import http.client
import os
import shutil
import time
import urllib.parse
import xml.etree.ElementTree as ET
from pathlib import Path
from typing import Any, Iterator

_EMPTY_SHA = hashlib.sha256(b"").hexdigest()


class ObjectStoreError(RuntimeError):
    def __init__(self, msg: str, status: int = 0) -> None:
        super().__init__(msg)
        self.status = status


class ObjectStore:
    """put / get / exists / list / delete on string keys; files move through ``upload`` / ``download``."""

    def put(self, key: str, data: bytes) -> None:
        raise NotImplementedError

    def get(self, key: str) -> bytes:
        raise NotImplementedError

    def exists(self, key: str) -> bool:
        raise NotImplementedError

    def list(self, prefix: str = "") -> list[str]:
        raise NotImplementedError

    def delete(self, key: str) -> None:
        raise NotImplementedError

    def upload(self, key: str, path: str | os.PathLike) -> None:
        self.put(key, Path(path).read_bytes())

    def download(self, key: str, path: str | os.PathLike) -> Path:
        p = Path(path)
        p.parent.mkdir(parents=True, exist_ok=True)
        tmp = p.with_name(f"{p.name}.{os.getpid()}.part")     # several ranks of one host may fetch the same object
        tmp.write_bytes(self.get(key))
        os.replace(tmp, p)
        return p

    def delete_prefix(self, prefix: str) -> int:
        keys = self.list(prefix)
        for k in keys:
            self.delete(k)
        return len(keys)


class DirObjectStore(ObjectStore):
    def __init__(self, root: str | os.PathLike, create: bool = True) -> None:
        self.root = Path(root)
        if create:
            self.root.mkdir(parents=True, exist_ok=True)

    def _p(self, key: str) -> Path:
        p = (self.root / key).resolve()
        if self.root.resolve() not in p.parents and p != self.root.resolve():
            raise ObjectStoreError(f"key escapes the store: {key}")
        return p

    def put(self, key: str, data: bytes) -> None:
        p = self._p(key)
        p.parent.mkdir(parents=True, exist_ok=True)
        tmp = p.with_name(p.name + ".part")
        tmp.write_bytes(data)
        os.replace(tmp, p)

    def get(self, key: str) -> bytes:
        try:
            return self._p(key).read_bytes()
        except FileNotFoundError:
            raise ObjectStoreError(f"no such key: {key}", 404) from None

    def exists(self, key: str) -> bool:
        return self._p(key).is_file()

    def list(self, prefix: str = "") -> list[str]:
        root = self.root.resolve()
        # walk only the directory the prefix points into (a prefix is "dir/dir/partial-name")
        start = (root / prefix).parent if prefix and not prefix.endswith("/") else root / prefix
        if not start.is_dir():
            return []
        return sorted(k for k in (str(p.relative_to(root)) for p in start.rglob("*") if p.is_file()) if k.startswith(prefix))

    def delete(self, key: str) -> None:
        p = self._p(key)
        if p.is_file():
            p.unlink()

    def upload(self, key: str, path: str | os.PathLike) -> None:
        p = self._p(key)
        p.parent.mkdir(parents=True, exist_ok=True)
        tmp = p.with_name(p.name + ".part")
        shutil.copyfile(path, tmp)
        os.replace(tmp, p)


# ------------------------------------------------------------------------------------------------------------------ S3
def _hmac(key: bytes, msg: str) -> bytes:
    return hmac.new(key, msg.encode(), hashlib.sha256).digest()


def _quote(s: str, safe: str = "-_.~") -> str:
    return urllib.parse.quote(s, safe=safe)


def sigv4_headers(method: str, host: str, path: str, query: dict[str, str], headers: dict[str, str], payload_sha256: str, *,
                  access_key: str, secret_key: str, region: str, service: str = "s3", now: _dt.datetime | None = None,
                  session_token: str | None = None) -> dict[str, str]:
    """Headers of one signed request (AWS Signature Version 4, header form). ``path`` is the already URI-encoded absolute path."""
    now = now or _dt.datetime.now(_dt.timezone.utc)
    amz_date, date = now.strftime("%Y%m%dT%H%M%SZ"), now.strftime("%Y%m%d")
    h = {k.lower(): " ".join(str(v).split()) for k, v in headers.items()}
    h.update({"host": host, "x-amz-date": amz_date, "x-amz-content-sha256": payload_sha256})
    if session_token:
        h["x-amz-security-token"] = session_token
    signed = ";".join(sorted(h))
    canonical_query = "&".join(f"{_quote(k)}={_quote(v)}" for k, v in sorted(query.items()))
    canonical = "\n".join([method, path, canonical_query, "".join(f"{k}:{h[k]}\n" for k in sorted(h)), signed, payload_sha256])
    scope = f"{date}/{region}/{service}/aws4_request"
    to_sign = "\n".join(["AWS4-HMAC-SHA256", amz_date, scope, hashlib.sha256(canonical.encode()).hexdigest()])
    k = _hmac(_hmac(_hmac(_hmac(("AWS4" + secret_key).encode(), date), region), service), "aws4_request")
    sig = hmac.new(k, to_sign.encode(), hashlib.sha256).hexdigest()
    h["authorization"] = f"AWS4-HMAC-SHA256 Credential={access_key}/{scope}, SignedHeaders={signed}, Signature={sig}"
    del h["host"]       # http.client writes it
    return h


class S3ObjectStore(ObjectStore):
    """``bucket`` (+ ``prefix``) on an S3-compatible endpoint."""

    def __init__(self, bucket: str, *, endpoint_url: str, access_key: str, secret_key: str, region: str = "us-east-1", prefix: str = "",
                 num_attempts: int = 3, connect_timeout: float = 60.0, read_timeout: float = 3600.0, session_token: str | None = None,
                 part_size: int = 64 << 20, multipart_threshold: int = 128 << 20, max_concurrency: int = 8,
                 ca_bundle: str | None = None) -> None:
        u = urllib.parse.urlparse(endpoint_url)
        if u.scheme not in ("http", "https") or not u.netloc:
            raise ValueError(f"endpoint_url must be http(s)://host[:port], got {endpoint_url!r}")
        self.scheme, self.host, self.base_path = u.scheme, u.netloc, u.path.rstrip("/")
        self.bucket, self.prefix = bucket, prefix.strip("/")
        self.access_key, self.secret_key, self.region, self.session_token = access_key, secret_key, region, session_token
        self.num_attempts = max(1, int(num_attempts))
        self.connect_timeout, self.read_timeout = float(connect_timeout), float(read_timeout)
        self.part_size, self.multipart_threshold = int(part_size), int(multipart_threshold)
        self.max_concurrency = max(1, int(max_concurrency))     # parts of one large object in flight (upload and ranged download)
        self._ssl: Any = None
        if self.scheme == "https":
            import ssl

            self._ssl = ssl.create_default_context(cafile=ca_bundle) if ca_bundle else ssl.create_default_context()   # e.g. a private MinIO CA

    # -- plumbing ------------------------------------------------------------------------------------------------
    def _key(self, key: str) -> str:
        return f"{self.prefix}/{key}" if self.prefix else key

    def _path(self, key: str | None) -> str:
        p = f"{self.base_path}/{_quote(self.bucket)}"
        if key is not None:
            p += "/" + _quote(self._key(key), safe="-_.~/")
        return p

    def _request(self, method: str, key: str | None, *, query: dict[str, str] | None = None, body: bytes = b"",
                 headers: dict[str, str] | None = None, ok: tuple[int, ...] = (200,), stream_to: Any = None) -> tuple[int, dict[str, str], bytes]:
        query = dict(query or {})
        path = self._path(key)
        sha = hashlib.sha256(body).hexdigest() if body else _EMPTY_SHA
        url = path + ("?" + "&".join(f"{_quote(k)}={_quote(v)}" for k, v in sorted(query.items())) if query else "")
        last: Exception | None = None
        for attempt in range(self.num_attempts):
            conn = None
            try:
                hdr = sigv4_headers(method, self.host, path, query, dict(headers or {}), sha, access_key=self.access_key,
                                    secret_key=self.secret_key, region=self.region, session_token=self.session_token)
                conn = (http.client.HTTPSConnection(self.host, timeout=self.connect_timeout, context=self._ssl) if self.scheme == "https"
                        else http.client.HTTPConnection(self.host, timeout=self.connect_timeout))
                conn.connect()
                conn.sock.settimeout(self.read_timeout)
                if body:
                    hdr["content-length"] = str(len(body))
                conn.request(method, url, body=body or None, headers=hdr)
                r = conn.getresponse()
                if stream_to is not None and r.status in ok:
                    while True:
                        chunk = r.read(8 << 20)
                        if not chunk:
                            break
                        stream_to.write(chunk)
                    data = b""
                else:
                    data = r.read()
                rh = {k.lower(): v for k, v in r.getheaders()}
                if r.status in ok:
                    return r.status, rh, data
                if r.status >= 500 or r.status == 429:         # transient: retry with back-off
                    last = ObjectStoreError(f"{method} {url}: HTTP {r.status} {data[:200]!r}", r.status)
                else:
                    raise ObjectStoreError(f"{method} {url}: HTTP {r.status} {data[:300]!r}", r.status)
            except (OSError, http.client.HTTPException) as e:
                last = e
            finally:
                if conn is not None:
                    conn.close()
            if attempt + 1 < self.num_attempts:
                time.sleep(min(5.0, 0.2 * 2 ** attempt))
        raise ObjectStoreError(f"{method} {url}: giving up after {self.num_attempts} attempts ({last})",
                               getattr(last, "status", 0)) from last

    # -- interface -----------------------------------------------------------------------------------------------
    def put(self, key: str, data: bytes) -> None:
        self._request("PUT", key, body=data)

    def get(self, key: str) -> bytes:
        try:
            return self._request("GET", key)[2]
        except ObjectStoreError as e:
            if e.status == 404:
                raise ObjectStoreError(f"no such key: {key}", 404) from None
            raise

    def exists(self, key: str) -> bool:
        try:
            self._request("HEAD", key)
            return True
        except ObjectStoreError as e:
            if e.status == 404:
                return False
            raise

    def delete(self, key: str) -> None:
        self._request("DELETE", key, ok=(200, 204))

    def _iter_list(self, prefix: str) -> Iterator[str]:
        token: str | None = None
        full = self._key(prefix) if (prefix or self.prefix) else ""
        if self.prefix and not prefix:
            full = self.prefix + "/"
        while True:
            q = {"list-type": "2", "prefix": full}
            if token:
                q["continuation-token"] = token
            _, _, data = self._request("GET", None, query=q)
            root = ET.fromstring(data)
            ns = root.tag[: root.tag.index("}") + 1] if root.tag.startswith("{") else ""
            for c in root.findall(f"{ns}Contents"):
                k = c.findtext(f"{ns}Key") or ""
                yield k[len(self.prefix) + 1:] if self.prefix else k
            if (root.findtext(f"{ns}IsTruncated") or "false").lower() != "true":
                return
            token = root.findtext(f"{ns}NextContinuationToken")
            if not token:
                return

    def list(self, prefix: str = "") -> list[str]:
        return sorted(self._iter_list(prefix))

    def upload(self, key: str, path: str | os.PathLike) -> None:
        p = Path(path)
        size = p.stat().st_size
        if size < self.multipart_threshold:
            self.put(key, p.read_bytes())
            return
        _, _, data = self._request("POST", key, query={"uploads": ""})
        root = ET.fromstring(data)
        ns = root.tag[: root.tag.index("}") + 1] if root.tag.startswith("{") else ""
        upload_id = root.findtext(f"{ns}UploadId")
        if not upload_id:
            raise ObjectStoreError(f"multipart upload of {key}: no UploadId in the response")
        n_parts = -(-size // self.part_size)

        def send(i: int) -> str:          # part i: its own file handle (positioned reads), its own retries inside _request
            with open(p, "rb") as f:
                f.seek(i * self.part_size)
                part = f.read(self.part_size)
            return self._request("PUT", key, query={"partNumber": str(i + 1), "uploadId": upload_id}, body=part)[1].get("etag", "")

        try:
            if self.max_concurrency > 1 and n_parts > 1:
                from concurrent.futures import ThreadPoolExecutor

                with ThreadPoolExecutor(max_workers=min(self.max_concurrency, n_parts), thread_name_prefix="s3-part") as ex:
                    etags = list(ex.map(send, range(n_parts)))
            else:
                etags = [send(i) for i in range(n_parts)]
            body = ("<CompleteMultipartUpload>" + "".join(f"<Part><PartNumber>{i + 1}</PartNumber><ETag>{e}</ETag></Part>"
                                                           for i, e in enumerate(etags)) + "</CompleteMultipartUpload>").encode()
            self._request("POST", key, query={"uploadId": upload_id}, body=body)
        except Exception:
            try:
                self._request("DELETE", key, query={"uploadId": upload_id}, ok=(200, 204))
            except ObjectStoreError:
                pass
            raise

    def size(self, key: str) -> int:
        try:
            return int(self._request("HEAD", key)[1].get("content-length", "0"))
        except ObjectStoreError as e:
            if e.status == 404:
                raise ObjectStoreError(f"no such key: {key}", 404) from None
            raise

    def download(self, key: str, path: str | os.PathLike) -> Path:
        """Stream the object to ``path``; objects above the multipart threshold come as parallel ranged GETs written at their offsets."""
        p = Path(path)
        p.parent.mkdir(parents=True, exist_ok=True)
        tmp = p.with_name(f"{p.name}.{os.getpid()}.part")
        try:
            total = self.size(key) if self.max_concurrency > 1 else 0
            if total >= self.multipart_threshold:
                from concurrent.futures import ThreadPoolExecutor

                with open(tmp, "wb") as f:
                    f.truncate(total)
                ranges = [(o, min(total, o + self.part_size) - 1) for o in range(0, total, self.part_size)]

                def fetch(r: tuple[int, int]) -> None:
                    _, _, data = self._request("GET", key, headers={"Range": f"bytes={r[0]}-{r[1]}"}, ok=(200, 206))
                    if len(data) != r[1] - r[0] + 1:
                        raise ObjectStoreError(f"GET {key} bytes={r[0]}-{r[1]}: got {len(data)} bytes")
                    with open(tmp, "r+b") as f:
                        f.seek(r[0])
                        f.write(data)

                with ThreadPoolExecutor(max_workers=min(self.max_concurrency, len(ranges)), thread_name_prefix="s3-range") as ex:
                    list(ex.map(fetch, ranges))
            else:
                with open(tmp, "wb") as f:
                    self._request("GET", key, stream_to=f)
        except ObjectStoreError as e:
            tmp.unlink(missing_ok=True)
            if e.status == 404:
                raise ObjectStoreError(f"no such key: {key}", 404) from None
            raise
        os.replace(tmp, p)
        return p


def remote_store_from_cfg(cfg: Any, env: dict[str, str] | None = None) -> ObjectStore | None:
    """The S3 endpoint of this run, or None (directory stand-in). A real store is used when an endpoint is configured —
    ``s3_comm_config.backend_kwargs.endpoint_url`` or ``S3_ENDPOINT_URL`` (the variable Composer's S3 backend reads; AWS itself:
    ``https://s3.<region>.amazonaws.com``) — and credentials are present (``AWS_ACCESS_KEY_ID`` / ``AWS_SECRET_ACCESS_KEY``). ``AWS_CA_BUNDLE`` (or
    ``backend_kwargs.verify``) names the CA file of an https endpoint with a private certificate."""
    env = dict(os.environ if env is None else env)
    sc = dict(cfg.get("s3_comm_config") or {})
    bk = dict(sc.get("backend_kwargs") or {})
    endpoint = bk.get("endpoint_url") or env.get("S3_ENDPOINT_URL")
    ak, sk = env.get("AWS_ACCESS_KEY_ID"), env.get("AWS_SECRET_ACCESS_KEY")
    if not endpoint:
        return None
    if not (ak and sk):
        raise ObjectStoreError("an S3 endpoint is configured but AWS_ACCESS_KEY_ID / AWS_SECRET_ACCESS_KEY are not set")
    cc = dict(bk.get("client_config") or {})
    return S3ObjectStore(str(sc.get("bucket_name", "checkpoints")), endpoint_url=str(endpoint), access_key=ak, secret_key=sk,
                         region=str(bk.get("region_name") or env.get("AWS_DEFAULT_REGION") or env.get("AWS_REGION") or "us-east-1"),
                         prefix=str(bk.get("prefix", "") or ""), num_attempts=int(sc.get("num_attempts", 3) or 3),
                         connect_timeout=float(cc.get("connect_timeout", 60) or 60), read_timeout=float(cc.get("read_timeout", 3600) or 3600),
                         session_token=env.get("AWS_SESSION_TOKEN") or None, ca_bundle=bk.get("verify") or env.get("AWS_CA_BUNDLE") or None)
