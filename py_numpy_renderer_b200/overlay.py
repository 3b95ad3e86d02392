"""Debug-frustum overlay (reference: obj/frustums.py:46-103 `draw_view_frustum`, obj/line.py:6-16
`bresenham_line`, called unconditionally at core.py:638).

The reference draws the edges of the *debug* camera's frustum, clipped to the camera's, as red lines with a depth
test, dashing the faces that look away, and blends the four neighbours of every line sample -- a short, strictly
sequential, order-dependent pass over at most 6 faces x 10 edges.  It is host work here as well: `Scene.render()`
first asks `segments()` whether anything would be drawn at all (nothing is when the debug frustum contains the
camera frustum, which is how every throughput scene is set up); only then does it fetch the float frame and the
z-buffer from the device and replay the pass with the same NumPy operations, duplicates and all.
"""
import numpy as np

from .transformation import extract_frustum_planes

# clip-space cube corners and its six faces (frustums.py:23-43)
_CORNERS = np.array([[-1.0, -1.0, 1.0, 1.0], [1.0, -1.0, 1.0, 1.0], [-1.0, 1.0, 1.0, 1.0], [1.0, 1.0, 1.0, 1.0],
                     [-1.0, 1.0, -1.0, 1.0], [1.0, 1.0, -1.0, 1.0], [-1.0, -1.0, -1.0, 1.0], [1.0, -1.0, -1.0, 1.0]])
_FACES = np.array([(2, 4, 5, 3), (0, 1, 7, 6), (0, 2, 3, 1), (5, 4, 6, 7), (3, 5, 7, 1), (4, 2, 0, 6)])


def _clip_polygon(poly, planes):
    """Sutherland-Hodgman against the camera planes (plane_intersection.py:59-86), same visit order and the same
    `line_plane_intersection(next, current)` argument order."""
    out = list(poly)
    for plane in planes:
        kept = []
        n = len(out)
        for i in range(n):
            cur, nxt = out[i], out[(i + 1) % n]
            cur_in, nxt_in = plane @ cur >= 0, plane @ nxt >= 0
            if cur_in:
                kept.append(cur)
            if cur_in ^ nxt_in:
                direction = cur - nxt
                den = plane @ direction
                if abs(den) >= 1e-10:
                    w = -(plane @ nxt) / den
                    if 0 <= w <= 1:
                        kept.append(nxt + w * direction)
        out = kept
    return np.array(out)


def _line_samples(start, end):
    """`bresenham_line`: a DDA from the right-hand end, int(steps) samples, last point excluded (line.py:6-16)."""
    delta = end - start
    if delta[0] > 0:
        return _line_samples(end, start)
    steps = max(abs(delta[:2]))
    if steps == 0:
        return start[None]
    return start + np.arange(int(steps))[:, None] * (delta / steps)


def camera_frustum_inside_debug_frustum(camera, debug_camera, margin=1e-7) -> bool:
    """True when all eight corners of the camera frustum lie strictly inside the six planes of the debug frustum.
    Then every face of the debug frustum -- a polygon on the boundary of a convex set that contains the camera frustum in
    its interior -- misses the camera frustum, and Sutherland-Hodgman against a convex region returns the empty polygon
    for an empty intersection: `segments_full` would return [].  8 x 6 dot products instead of 36 Python-level clip steps
    (0.8 ms per `Scene.render()` call on the host, about half of a call)."""
    with np.errstate(all='ignore'):
        corners = _CORNERS @ np.linalg.inv(camera.MVP)
        corners = corners / corners[:, [3]]
        return bool((corners @ extract_frustum_planes(debug_camera.MVP).T > margin).all())


def segments(camera, debug_camera):
    """Projected, clipped faces of the debug frustum: list of (polygon (k,4) with linearised z, dashed flag).
    Empty list = the overlay touches no pixel."""
    try:
        if camera_frustum_inside_debug_frustum(camera, debug_camera):
            return []
    except np.linalg.LinAlgError:
        pass
    return segments_full(camera, debug_camera)


def segments_full(camera, debug_camera):
    """The reference's own sequence (frustums.py:46-75): clip every face of the debug frustum to the camera frustum."""
    world = _CORNERS @ np.linalg.inv(debug_camera.MVP)
    world /= world[:, [3]]
    planes = extract_frustum_planes(camera.MVP)
    probe = np.append(camera.position, 1) @ debug_camera.MVP
    inside = bool(-probe[3] < probe[0] < probe[3] and -probe[3] < probe[1] < probe[3] and -probe[3] < probe[2] < probe[3])
    near, far = camera.near, camera.far
    out = []
    for face in world[_FACES]:
        face = _clip_polygon(face, planes)
        if face.shape[0] < 3:
            continue
        face = face @ camera.MVP
        face /= face[:, [3]]
        face = face @ camera.viewport
        a, b, c = face[0, :3], face[1, :3], face[2, :3]
        normal = np.cross(b - a, c - a)
        face[:, 2] = (2 * near * far) / (far + near - face[:, 2] * (far - near))
        out.append((face, bool(normal[2] > 0 and not inside)))
    return out


def apply(frame, z_buffer, camera, debug_camera, sign, polygons=None):
    """Replay frustums.py:76-103 on `frame` (float32 (H,W,3), buffer rows) and `z_buffer` (float64 (H,W)) in place."""
    if polygons is None:
        polygons = segments(camera, debug_camera)
    color = np.array((1., 0., 0.))
    clip_x, clip_y = np.array(frame.shape[:2]) - 1
    for face, dashed in polygons:
        count = len(face)
        for i in range(count):
            pxls = _line_samples(face[i], face[(i + 1) % count])
            if dashed:
                keep = np.bitwise_and(np.arange(len(pxls)) // 13, 1, dtype=np.int8).view(np.bool_)
                pxls = pxls[keep]
            y, x, z, _ = pxls.T
            x = x.astype(np.int32) - 1
            y = y.astype(np.int32) - 1
            visible = ((z_buffer[x, y] - z) * sign >= 0)
            x, y, z = x[visible], y[visible], z[visible]
            z_buffer[x, y] = z
            frame[x, y] = color
            for step in (-1, 1):
                xs, ys = np.clip(x + step, a_min=0, a_max=clip_x), np.clip(y + step, a_min=0, a_max=clip_y)
                z_buffer[xs, y] = z
                z_buffer[x, ys] = z
                frame[xs, y] = frame[xs, y] * 0.5 + color / 2
                frame[x, ys] = frame[x, ys] * 0.5 + color / 2
    return frame


def tonemap(frame):
    """core.py:640 on the host (only used when the overlay forced the frame back to float)."""
    return (frame[::-1] ** 0.8 * 255).astype(np.uint8)
