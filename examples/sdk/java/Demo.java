// Java client demo for the deeprec_b200 model server (JDK 11+, no dependencies):
//   javac Demo.java && java Demo http://127.0.0.1:8500 ctr
import java.net.URI;
import java.net.http.HttpClient;
import java.net.http.HttpRequest;
import java.net.http.HttpResponse;
import java.util.Random;

public class Demo {
    public static void main(String[] args) throws Exception {
        String base = args.length > 0 ? args[0] : "http://127.0.0.1:8500";
        String model = args.length > 1 ? args[1] : "ctr";
        int batch = 4;
        Random rng = new Random(0);
        StringBuilder body = new StringBuilder("{\"dense\": [");
        for (int i = 0; i < batch; ++i) {                       // [B][13] floats
            body.append(i == 0 ? "[" : ", [");
            for (int k = 0; k < 13; ++k) body.append(k == 0 ? "" : ", ").append(rng.nextFloat() * 2 - 1);
            body.append("]");
        }
        body.append("], \"ids\": [");
        for (int t = 0; t < 26; ++t) {                          // [26][B] ids
            body.append(t == 0 ? "[" : ", [");
            for (int i = 0; i < batch; ++i) body.append(i == 0 ? "" : ", ").append(rng.nextInt(1000));
            body.append("]");
        }
        body.append("]}");
        HttpRequest req = HttpRequest.newBuilder(URI.create(base + "/v1/models/" + model + ":predict"))
                .header("Content-Type", "application/json")
                .POST(HttpRequest.BodyPublishers.ofString(body.toString()))
                .build();
        HttpResponse<String> resp = HttpClient.newHttpClient().send(req, HttpResponse.BodyHandlers.ofString());
        if (resp.statusCode() != 200) {
            System.err.println("HTTP " + resp.statusCode() + ": " + resp.body());
            System.exit(1);
        }
        System.out.println(resp.body());                        // {"predictions": [...], "model_version": v}
    }
}
